#!/usr/bin/env python
"""Small end-to-end run of every kernel of libdpcu.so, sized for compute-sanitizer (memcheck / racecheck / synccheck /
initcheck are 10-100x slower than a plain run).  Every result is still compared with the oracle.

    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.loader import Port  # noqa: E402
from pipeline_b200 import capi, scenes  # noqa: E402
from tests import cases  # noqa: E402


def main():
    port = Port()
    n = 40000 + 77
    lower4, extent4, upper4, mats, tidx = cases.random_case(n, seed=scenes.SEED_C4)
    cams = np.ascontiguousarray(np.concatenate([scenes.cube_map_cameras((30.0, 5.0, -20.0)), [scenes.camera_c2(), scenes.orbit_camera(3)]]), np.float32)
    launches = 0
    for nv in (1, 2, 6, 8):
        want = [port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), cams[v]) for v in range(nv)]
        # (kernel form, list built in the line-granular kernel?, DPCU_CULL_OPT_LIST_OFFSETS of the one thread per object forms)
        for kernel, fuse_list, offsets in ((0, 1, 0), (1, 1, 0), (1, 1, 1), (1, 1, 2), (2, 1, 0), (3, 1, 0), (3, 1, 1), (3, 1, 2), (4, 1, 0),
                                           (4, 0, 0), (7, 1, 0), (7, 0, 0), (8, 1, 0)):
            ctx = capi.Cull(0)
            ctx.set_option(capi.OPT_KERNEL, kernel)
            ctx.set_option(capi.OPT_FUSE_LIST, fuse_list)
            ctx.set_option(capi.OPT_LIST_OFFSETS, offsets)
            ctx.set_objects(lower4, extent4, tidx)
            ctx.set_matrices(mats.reshape(-1))
            res = [ctx.result_create() for _ in range(nv)]
            mirrors = []
            for r in res:
                hb = [capi.HostBuffer(((n + 31) // 32) * 4), capi.HostBuffer(n * 4), capi.HostBuffer(4)]
                mirrors.append(hb)
                r.set_host_mirror(hb[0].array(np.uint32), hb[1].array(np.uint32), hb[2].array(np.uint32))
            for frame in range(2):
                ctx.run(res, cams[:nv] if frame == 0 else cams[1:nv + 1] if nv < 8 else cams[:nv])
                for r in res:
                    r.synchronize()
            for v in range(nv):
                ref = want[v] if nv == 8 else port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), cams[1 + v])
                assert np.array_equal(res[v].bits(), ref), (nv, kernel, v)
                assert np.array_equal(mirrors[v][0].array(np.uint32)[:len(ref)], ref)
                cnt = int(mirrors[v][2].array(np.uint32)[0])
                assert np.array_equal(mirrors[v][1].array(np.uint32)[:cnt], res[v].changed())
            res[0].build_visible_list()
            assert len(res[0].visible()) == int(np.unpackbits(res[0].bits().view(np.uint8)).sum())
            ctx.bounding_box()
            launches += ctx.launches()
            for r in res:
                r.set_host_mirror(None, None, None)
                r.close()
            ctx.close()
            for hb in mirrors:
                for b in hb:
                    b.close()
    # batched edits, bit moves
    ctx = capi.Cull(0)
    ctx.set_objects(lower4[:30000], extent4[:30000], tidx[:30000] % 30000)
    ctx.set_matrices(mats.reshape(-1))
    r = ctx.result_create()
    ctx.run([r], cams[6])
    ctx.set_object_count(31000)
    idx = np.arange(29000, 31000, dtype=np.uint32)
    ctx.update_objects(idx, lower4[idx], extent4[idx], tidx[idx] % 30000)
    r.move_bit(5, 7)
    r.update_words(np.array([3], np.uint32), np.array([0xF0F0F0F0], np.uint32))
    ctx.run([r], cams[7])
    launches += ctx.launches()
    r.close(), ctx.close()
    # transform tree: narrow and wide K1, fused leaf level with and without a peer gather, dirty-only refresh
    levels = (8, 64, 512, 65536)
    entries, offsets, n_nodes = scenes.hierarchy_topology(levels)
    local = np.zeros((n_nodes, 4, 4), np.float32)
    local[0] = np.eye(4, dtype=np.float32)
    local[1:] = scenes.hierarchy_locals(scenes.SEED_C3, 1, n_nodes - 1, frame=1)
    t = capi.Tree(0)
    t.set_topology(entries, offsets, n_nodes)
    t.set_locals(0, local)
    nl = levels[-1]
    lo, ex, _, _, _ = cases.random_case(nl)
    ctx = capi.Cull(0)
    ctx.set_objects(lo, ex, np.arange(n_nodes - nl, n_nodes, dtype=np.uint32))
    r = ctx.result_create()
    full = capi.Buffer(((nl + 31) // 32) * 4 + 4096)
    full.fill(0)
    for peers in (0, 1):
        if peers:
            r.set_peer_bits([full.ptr], 0)
        t.mark_dirty(1, n_nodes - 1)
        ctx.run_with_tree(t, [r], cams[6])
    world = np.zeros_like(local)
    world[0] = local[0]
    nw = (n_nodes + 31) // 32
    port.tree_compute(local, world, entries, offsets, np.full(nw, 0xFFFFFFFF, np.uint32), np.zeros(nw, np.uint32))
    assert np.array_equal(t.world().view(np.uint32), world.view(np.uint32))
    want = port.cull_bits(lo, ex, np.arange(n_nodes - nl, n_nodes, dtype=np.uint32), world.reshape(-1), cams[6])
    assert np.array_equal(r.bits(), want)
    got = full.download(np.zeros((nl + 31) // 32, np.uint32))
    assert np.array_equal(got, want)
    host = np.zeros_like(local)
    t.update_locals(np.array([5, 900], np.uint32), local[[5, 900]])
    t.compute()
    t.refresh_host_world(host)
    launches += ctx.launches() + t.launches()
    r.close(), ctx.close(), t.close(), full.close()
    print("sanitize smoke ok: %d kernel launches, every result equal to the oracle" % launches)


if __name__ == "__main__":
    main()
