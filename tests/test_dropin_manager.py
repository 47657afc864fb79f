"""The drop-in test: ONE driver (oracle/ref_shim.cpp, the calls dp::sg::xbar::culling::CullingImpl
makes) runs the reference's dp::culling::cpu::Manager and the new dp::culling::cuda::Manager through
the same virtual dp::culling::Manager API; their answers must be identical.

tests/cpp/_build/libdpdropin.so is built in the build container (it compiles the new C++ host layer
against the reference headers where they lie) and travels to the GPU box."""
import os

import numpy as np
import pytest

from pipeline_b200 import scenes
from tests import cases
from tests.engines import RefEngine, run_lifecycle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "tests", "cpp", "_build", "libdpdropin.so")


@pytest.fixture(scope="module")
def dropin():
    from oracle.loader import Reference
    if os.path.isdir("/root/reference"):
        import subprocess
        import __graft_entry__ as g
        g.build_product()
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    if not os.path.exists(DROPIN):
        pytest.skip("tests/cpp/_build/libdpdropin.so not present (built only where /root/reference exists)")
    return Reference(DROPIN)


def test_dropin_library_fails_loudly_without_gpu(dropin):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cpu = dropin.cull(0)           # the reference backend works anywhere
    cpu.close()
    with pytest.raises(RuntimeError) as e:
        dropin.cull(1)             # dp::culling::cuda::Manager::create() throws: no silent CPU path
    assert "no CPU fallback" in str(e.value)


@pytest.mark.gpu
def test_same_driver_two_backends_random_frames(dropin):
    lower4, extent4, upper4, mats, tidx = cases.random_case(50000)
    a, b = RefEngine(dropin, 0), RefEngine(dropin, 1)
    for e in (a, b):
        e.add(lower4, upper4, tidx)
        e.set_matrices(mats.reshape(-1))
    for vp in cases.frames(5):
        ab, ac = a.cull(vp)
        bb, bc = b.cull(vp)
        assert np.array_equal(ab, bb)
        assert np.array_equal(ac, bc)
    assert np.array_equal(a.bounding_box(), b.bounding_box())
    a.close(), b.close()


@pytest.mark.gpu
def test_same_driver_two_backends_lifecycle(dropin, golden):
    g = golden("lifecycle")
    lower4, extent4, upper4, mats, tidx = cases.random_case(5000, seed=0x11FE)
    res = run_lifecycle(RefEngine(dropin, 1), cases.lifecycle_script(), lower4, upper4, mats, cases.frames())
    for k, (bits, changed, n) in enumerate(res):
        assert n == int(g["count%d" % k])
        assert np.array_equal(bits, g["bits%d" % k]), "step %d bits" % k
        assert np.array_equal(changed, g["changed%d" % k]), "step %d changed" % k


@pytest.mark.gpu
def test_same_driver_two_backends_special_and_gather(dropin, golden):
    g = golden("special4k")
    lower4, extent4, upper4, mats, tidx, vps = cases.special_case()
    e = RefEngine(dropin, 1)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    for k, vp in enumerate(vps):
        bits, changed = e.cull(vp)
        assert np.array_equal(bits, g["bits%d" % k])
        assert np.array_equal(changed, g["changed%d" % k])
    e.close()
    g = golden("gather5k")
    lower4, extent4, upper4, raw, tidx, stride = cases.gather_case()
    e = RefEngine(dropin, 1)
    e.add(lower4, upper4, tidx)
    e.set_matrices(raw.reshape(-1), stride, len(raw))
    bits, changed = e.cull(scenes.camera_c2())
    assert np.array_equal(bits, g["bits"]) and np.array_equal(changed, g["changed"])
    assert np.array_equal(e.bounding_box(), g["bbox"])
    e.close()


@pytest.mark.gpu
def test_dirty_matrix_protocol(dropin):
    """groupMatrixChanged -> only the flagged matrices are re-read (GroupBitSet.h:140-150)."""
    lower4, extent4, upper4, mats, tidx = cases.random_case(4096)
    mats = mats.copy()
    a, b = RefEngine(dropin, 0), RefEngine(dropin, 1)
    for e in (a, b):
        e.add(lower4, upper4, tidx)
        e.set_matrices(mats.reshape(-1))
    vp = scenes.camera_c2()
    a.cull(vp), b.cull(vp)
    idx = np.arange(0, 4096, 5, dtype=np.uint32)
    mats[idx.astype(np.int64), 3, 2] -= np.float32(250.0)
    for e in (a, b):
        e.set_matrices(mats.reshape(-1))        # same pointer / count / stride: nothing is re-read by itself
        e.matrices_changed(idx)
    ab, ac = a.cull(vp)
    bb, bc = b.cull(vp)
    assert np.array_equal(ab, bb) and np.array_equal(ac, bc) and len(bc) > 0
    a.close(), b.close()


@pytest.mark.gpu
def test_live_object_edit_reaches_the_device(dropin, port):
    """objectSetBoundingBox / objectSetTransformIndex on an object that is already in the group."""
    lower4, extent4, upper4, mats, tidx = cases.random_case(2000)
    e = RefEngine(dropin, 1)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    vp = scenes.camera_c2()
    e.cull(vp)
    lower4, upper4, tidx = lower4.copy(), upper4.copy(), tidx.copy()
    for i in (3, 777, 1999):
        lower4[i, :3] -= 50.0
        upper4[i, :3] += 50.0
        tidx[i] = (i * 7) % 2000
        e.set_object(i, lower4[i], upper4[i], int(tidx[i]))
    bits, _ = e.cull(vp)
    ext = port.box_extent(np.ascontiguousarray(lower4), np.ascontiguousarray(upper4))
    want = port.cull_bits(np.ascontiguousarray(lower4), ext, tidx, mats.reshape(-1), vp)
    assert np.array_equal(bits, want)
    e.close()


@pytest.mark.gpu
def test_multi_view_and_device_matrices_extensions(dropin, port):
    from pipeline_b200 import capi
    n = 20000
    lower4, extent4, upper4, mats, tidx = cases.random_case(n, seed=scenes.SEED_C4)
    s = dropin.cull(1)
    s.add_objects(np.ascontiguousarray(lower4[:, :3]), np.ascontiguousarray(upper4[:, :3]), tidx)
    buf = capi.Buffer(mats.nbytes)
    buf.upload(mats)
    s.set_device_matrices(buf.ptr, n)
    results = [s.result_create() for _ in range(6)]
    vps = scenes.cube_map_cameras()
    s.cull_multi(results, vps)
    for v in range(6):
        want = port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), vps[v])
        assert np.array_equal(s.visible_bits(results[v]), want), v
    s.close()
    buf.close()
