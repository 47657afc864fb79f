"""The on-device synthetic-scene generator must replay the host generator bit for bit
("a host replay of any slice matches bit-for-bit")."""
import numpy as np
import pytest

from pipeline_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("first,count,base", [(0, 4099, 0), (123456789, 2048, 123456789), ((1 << 28) - 1000, 1000, 0)])
def test_device_generator_matches_host(first, count, base):
    from pipeline_b200 import capi
    lo, ex, mt = capi.Buffer(count * 16), capi.Buffer(count * 16), capi.Buffer(count * 64)
    capi.scene_generate(scenes.SEED_C4, first, count, base, lo.ptr, ex.ptr, mt.ptr)
    capi.device_sync()
    lower4, extent4, upper4, mats, tidx = scenes.random_objects(scenes.SEED_C4, first, count)
    got_lo = lo.download(np.zeros((count, 4), np.float32))
    got_ex = ex.download(np.zeros((count, 4), np.float32))
    got_mt = mt.download(np.zeros((count, 4, 4), np.float32))
    assert np.array_equal(got_lo[:, :3].view(np.uint32), lower4[:, :3].view(np.uint32))
    assert np.array_equal(got_lo[:, 3].view(np.uint32), (tidx - np.uint32(base)).astype(np.uint32))
    assert np.array_equal(got_ex.view(np.uint32), extent4.view(np.uint32))
    assert np.array_equal(got_mt.view(np.uint32), mats.view(np.uint32))
    for b in (lo, ex, mt):
        b.close()
