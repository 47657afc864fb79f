"""N > 1 host logic on CPU: two gloo ranks each cull their slice (with the oracle port standing in
for the GPU), the bitset slices are gathered, and the result must equal the unsharded cull."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from pipeline_b200 import sharding  # noqa: E402


def test_shard_ranges_cover_and_align():
    for n in (0, 1, 1023, 1024, 1025, 100003, 1 << 20, (1 << 28) + 5):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                first, count = sharding.shard_range(n, world, r)
                assert first == nxt and first % 1024 == 0 or first == n
                nxt = first + count
            assert nxt == n


def _worker(rank, world, port_file, n, tmpdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.loader import Port
    from pipeline_b200 import scenes, sharding
    dist.init_process_group("gloo", init_method="file://" + port_file, rank=rank, world_size=world)
    port = Port()
    first, count = sharding.shard_range(n, world, rank)
    lower4, extent4, upper4, mats, tidx = scenes.random_objects(scenes.SEED_C5, first, count)
    tidx = (tidx - np.uint32(first)).astype(np.uint32)
    res = port.result_resize(np.zeros(0, np.uint32), 0, count)
    outs = []
    for f in range(3):
        vp = scenes.orbit_camera(f)
        words = port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), vp)
        changed = port.update_changed(words, res, count)
        # gather: equal-sized word slices (pad the last one)
        per = sharding.total_words(sharding.shard_range(n, world, 0)[1])
        padded = np.zeros(per, np.uint32)
        padded[:len(words)] = words
        gathered = sharding.allgather_words(dist, padded, per)
        full = np.zeros(sharding.total_words(n) + per, np.uint32)
        for r, t in enumerate(gathered):
            fr, cr = sharding.shard_range(n, world, r)
            sharding.place_words(full, t.numpy().view(np.uint32)[:sharding.total_words(cr)], fr)
        lists = [None] * world
        dist.all_gather_object(lists, changed)
        firsts = [sharding.shard_range(n, world, r)[0] for r in range(world)]
        merged = sharding.merge_changed(lists, firsts)
        outs.append((full[:sharding.total_words(n)].copy(), merged))
    if rank == 0:
        np.savez(os.path.join(tmpdir, "out.npz"), **{"bits%d" % i: o[0] for i, o in enumerate(outs)},
                 **{"chg%d" % i: o[1] for i, o in enumerate(outs)})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_unsharded(tmp_path, port):
    import torch.multiprocessing as mp
    from pipeline_b200 import scenes
    n = 50000 + 777
    port_file = str(tmp_path / "rendezvous")
    mp.spawn(_worker, args=(2, port_file, n, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "out.npz")
    lower4, extent4, upper4, mats, tidx = scenes.random_objects(scenes.SEED_C5, 0, n)
    res = port.result_resize(np.zeros(0, np.uint32), 0, n)
    for f in range(3):
        want = port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), scenes.orbit_camera(f))
        want_changed = port.update_changed(want, res, n)
        assert np.array_equal(got["bits%d" % f], want)
        assert np.array_equal(got["chg%d" % f], want_changed.astype(np.uint64))
        assert np.all(np.diff(got["chg%d" % f].astype(np.int64)) > 0)


@pytest.mark.parametrize("world", [2, 4])
def test_tree_shards_equal_whole_tree(port, world):
    """C3 across GPUs: upper levels replicated, leaf level sharded with the objects.  Every rank's
    shard tree, propagated and culled on its own (the oracle standing in for the GPU), gives the world matrices and
    the bitset words of its slice of the whole tree."""
    from pipeline_b200 import scenes
    levels = (4, 16, 64, 4096)
    entries, offsets, n_nodes = scenes.hierarchy_topology(levels)
    local = np.zeros((n_nodes, 4, 4), np.float32)
    local[0] = np.eye(4, dtype=np.float32)
    local[1:] = scenes.hierarchy_locals(scenes.SEED_C3, 1, n_nodes - 1, frame=3)
    world_m = np.zeros_like(local)
    world_m[0] = local[0]
    nw = (n_nodes + 31) // 32
    port.tree_compute(local, world_m, entries, offsets, np.full(nw, 0xFFFFFFFF, np.uint32), np.zeros(nw, np.uint32))
    n = levels[-1]
    first_leaf = n_nodes - n
    lower4, extent4, upper4, _, _ = scenes.random_objects(scenes.SEED_C3, 0, n)
    lower4[:, :3] *= 0.2
    extent4[:, :3] *= 0.2
    vp = scenes.mat_mul(scenes.make_look_at((0, 0, 120), (0, 0, 0), (0, 1, 0)), scenes.make_perspective(50.0, 1.5, 1.0, 400.0))
    whole = port.cull_bits(lower4, extent4, np.arange(first_leaf, n_nodes, dtype=np.uint32), world_m.reshape(-1), vp)
    assert 0 < int(np.unpackbits(whole.view(np.uint8)).sum()) < n
    full = np.zeros(sharding.total_words(n), np.uint32)
    covered = 0
    for rank in range(world):
        e, o, nn, first_leaf_local, first, count, gmap = sharding.tree_shard(levels, world, rank)
        assert first == covered and nn == first_leaf_local + count and len(gmap) == nn
        covered += count
        assert len(o) == len(levels) + 1 and o[-1] - o[-2] == count
        loc = local[gmap.astype(np.int64)]                      # local matrices are a function of the global node index
        wm = np.zeros_like(loc)
        wm[0] = loc[0]
        nws = (nn + 31) // 32
        port.tree_compute(loc, wm, e, o, np.full(nws, 0xFFFFFFFF, np.uint32), np.zeros(nws, np.uint32))
        assert np.array_equal(wm[first_leaf_local:].view(np.uint32), world_m[first_leaf + first:first_leaf + first + count].view(np.uint32))
        tidx = np.arange(first_leaf_local, nn, dtype=np.uint32)
        words = port.cull_bits(lower4[first:first + count], extent4[first:first + count], tidx, wm.reshape(-1), vp)
        sharding.place_words(full, words, first)
    assert covered == n
    assert np.array_equal(full, whole)
