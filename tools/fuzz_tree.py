#!/usr/bin/env python
"""Differential fuzz of the transform path (GPU): random hierarchies (random fan-out, entries of a level in node order or
shuffled - the coalesced and the strided kernel paths), narrow / wide / default level kernels, random dirty sets over
several frames (everything, ranges, sparse, nothing), some non-affine / non-finite locals; world matrices (bitwise) and the
published dirty set against the oracle port's Tree::compute.  Every few rounds the last level also runs fused into a cull
(dpcuCullRunWithTree, objects bound to the leaves in entry order) and the bits are compared as well."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.loader import Port  # noqa: E402  (test infrastructure: the checker)
from pipeline_b200 import capi, scenes  # noqa: E402


def one_round(port, rng, log):
    n_levels = int(rng.randint(1, 5))
    sizes = [int(rng.randint(1, 40))]
    for _ in range(n_levels - 1):
        sizes.append(int(min(sizes[-1] * rng.randint(1, 40) + rng.randint(0, 5), 120000)))
    shuffle = bool(rng.randint(2))
    wide_min = [None, 0, 1 << 40][rng.randint(3)]
    log("levels=%s shuffled=%d wide_min=%s" % (sizes, shuffle, wide_min))
    entries, offsets = [], [0]
    first_prev, n_prev, nxt = 0, 1, 1
    for n in sizes:
        node = np.arange(nxt, nxt + n, dtype=np.uint32)
        parent = (first_prev + np.sort(rng.randint(0, n_prev, size=n))).astype(np.uint32)
        if shuffle:
            p = rng.permutation(n)
            node, parent = node[p], parent[p]
        entries.append(np.stack([parent, node], axis=1))
        offsets.append(offsets[-1] + n)
        first_prev, n_prev, nxt = nxt, n, nxt + n
    entries = np.ascontiguousarray(np.concatenate(entries), np.uint32)
    offsets = np.asarray(offsets, np.uint32)
    n_nodes = nxt
    local = np.zeros((n_nodes, 4, 4), np.float32)
    local[0] = np.eye(4, dtype=np.float32)
    local[1:] = scenes.hierarchy_locals(int(rng.randint(1, 1 << 30)), 1, n_nodes - 1, frame=0)
    if rng.randint(3) == 0 and n_nodes > 8:                  # a few nasty locals
        bad = rng.choice(np.arange(1, n_nodes), size=min(n_nodes // 8, 64), replace=False)
        vals = np.array([0.0, -1.0, 2.0, 1e-40, 1e30, np.inf, np.nan, 0.5], np.float32)
        local[bad] = rng.choice(vals, size=(len(bad), 4, 4))
    t = capi.Tree(0)
    if wide_min is not None:
        t.set_option(capi.TREE_OPT_WIDE_MIN_NODES, wide_min)
    t.set_topology(entries, offsets, n_nodes)
    t.set_locals(0, local)
    nw = (n_nodes + 31) // 32
    world = np.zeros_like(local)
    world[0] = np.eye(4, dtype=np.float32)
    dl = np.full(nw, 0xFFFFFFFF, np.uint32)
    dw = np.zeros(nw, np.uint32)
    tail = np.uint32((1 << (n_nodes % 32)) - 1) if n_nodes % 32 else np.uint32(0xFFFFFFFF)
    # a culling context bound to the leaves in entry order, for the fused form
    fused = bool(rng.randint(2)) and sizes[-1] >= 1
    if fused:
        leaves = entries[offsets[-2]:offsets[-1], 1].astype(np.uint32)
        nobj = len(leaves)
        lo4, ex4, up4, _, _ = scenes.random_objects(int(rng.randint(1, 1 << 30)), 0, nobj)
        ctx = capi.Cull(0)
        ctx.set_objects(np.ascontiguousarray(lo4), np.ascontiguousarray(ex4), leaves)
        ctx.set_option(capi.OPT_LIST_OFFSETS, int(rng.randint(0, 4)))
        res = ctx.result_create()
        state = port.result_resize(np.zeros(0, np.uint32), 0, nobj)
    checked = 0
    for frame in range(int(rng.randint(2, 6))):
        if frame:
            kind = rng.randint(4)
            upd = scenes.hierarchy_locals(int(rng.randint(1, 1 << 30)), 1, n_nodes - 1, frame=frame)
            if kind == 0:                                     # everything
                local[1:] = upd
                t.set_locals(1, upd)
                dl[:] = 0xFFFFFFFF
            elif kind == 1 and n_nodes > 2:                   # a contiguous range
                a = int(rng.randint(1, n_nodes - 1)); b = int(rng.randint(a + 1, n_nodes + 1))
                local[a:b] = upd[a - 1:b - 1]
                t.set_locals(a, local[a:b])
                idx = np.arange(a, b)
                np.bitwise_or.at(dl, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
            elif kind == 2 and n_nodes > 2:                   # a sparse batch
                k = int(rng.randint(1, max(2, n_nodes // 10)))
                idx = rng.choice(np.arange(1, n_nodes), size=min(k, n_nodes - 1), replace=False).astype(np.uint32)
                local[idx] = upd[idx - 1]
                t.update_locals(idx, local[idx])
                np.bitwise_or.at(dl, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
            # kind == 3: nothing dirty
        vp = scenes.orbit_camera(int(rng.randint(40)))
        if fused:
            ctx.run_with_tree(t, [res], np.ascontiguousarray(vp, np.float32).reshape(1, 16))
        else:
            t.compute()
        dw[:] = 0
        port.tree_compute(local, world, entries, offsets, dl, dw)
        got_w = t.world()
        # bit for bit, except the payload of a NaN: x86 propagates / produces 0xffc00000-style NaNs, the GPU its canonical
        # 0x7fffffff - a NaN must be a NaN in the same place, every other value must have the same bits
        gu, wu = got_w.view(np.uint32), world.view(np.uint32)
        same = (gu == wu) | (np.isnan(got_w) & np.isnan(world))
        assert same.all(), "frame %d: %d world matrices differ" % (frame, int((~same).any(axis=(1, 2)).sum()))
        got = t.dirty_world()
        got[-1] &= tail
        assert np.array_equal(got, dw), "frame %d: dirty set differs" % frame
        if fused:
            want = port.cull_bits(np.ascontiguousarray(lo4), np.ascontiguousarray(ex4), leaves, world.reshape(-1), vp)
            assert np.array_equal(res.bits(), want), "frame %d: fused cull bits differ" % frame
            assert np.array_equal(res.changed(), port.update_changed(want, state, nobj)), "frame %d: fused cull changed list" % frame
        checked += n_nodes
    if fused:
        res.close()
        ctx.close()
    t.close()
    return checked


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=None)
    a = ap.parse_args()
    port = Port()
    master = np.random.RandomState(a.seed if a.seed is not None else int(time.time()) & 0x7FFFFFFF)
    t0 = time.time()
    rounds = nodes = 0
    while time.time() - t0 < a.seconds:
        seed = int(master.randint(1, 1 << 30))
        lines = []
        try:
            nodes += one_round(port, np.random.RandomState(seed), lines.append)
        except Exception:
            print("FAILED in round seed %d: %s" % (seed, "; ".join(lines)))
            raise
        rounds += 1
    print("tree fuzz ok: %d rounds, %d node propagations equal to the oracle, bit for bit (%.0f s)" % (rounds, nodes, time.time() - t0))


if __name__ == "__main__":
    main()
