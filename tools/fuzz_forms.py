#!/usr/bin/env python
"""Differential fuzz of the cull path through the C ABI (GPU): random group sizes, view counts, kernel forms, list-offset
modes, line sizes, host mirrors on / off, object-count changes and live edits between frames - every frame's bitsets and
ordered changed lists compared with the oracle port (tests/engines.py's bookkeeping).  `--seconds S` bounds the run; a
failure prints the seed and the step so that it can be replayed (`--seed`).  tests/test_cuda_parity.py runs a short slice."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.loader import Port  # noqa: E402  (test infrastructure: the checker)
from pipeline_b200 import capi, scenes  # noqa: E402
from tests import cases  # noqa: E402

FORMS = [(0, 1), (1, 1), (2, 1), (3, 1), (4, 1), (4, 0), (5, 1), (7, 1), (7, 0), (8, 1)]   # (DPCU_KERNEL_*, fuse list)


def one_round(port, rng, log):
    n = int(rng.choice([rng.randint(1, 300), rng.randint(300, 20000), rng.randint(20000, 400000)]))
    nv = int(rng.choice([1, 1, 2, 3, 6, 8]))
    kernel, fuse = FORMS[rng.randint(len(FORMS))]
    offsets = int(rng.randint(0, 4))
    line_words = int(rng.choice([0, 0, 8, 16, 32]))
    mirror = bool(rng.randint(2))
    seed = int(rng.randint(1, 1 << 30))
    log("n=%d views=%d kernel=%d fuse=%d offsets=%d line_words=%d mirror=%d scene=%d" % (n, nv, kernel, fuse, offsets, line_words, mirror, seed))
    cap = n + 5000
    lower4, extent4, upper4, mats, tidx = cases.random_case(cap, seed=seed)
    if rng.randint(3) == 0:                                   # shared / permuted transforms
        tidx = rng.randint(0, cap, size=cap).astype(np.uint32)
    if rng.randint(4) == 0:                                   # nasty values: non-affine rows, NaN / Inf, denormals, corners on planes
        sl, se, su, sm, st, svps = cases.special_case(min(cap, 4096), seed=seed & 0xFFFF)
        where = rng.choice(cap, size=len(sl), replace=False)
        mats = mats.copy()
        lower4, extent4 = lower4.copy(), extent4.copy()
        mats[where] = sm
        lower4[where, :3], extent4[where, :3] = sl[:, :3], se[:, :3]
        special_views = list(svps)
    else:
        special_views = []
    flat = mats.reshape(-1)
    cams = special_views + list(scenes.cube_map_cameras((float(rng.uniform(-40, 40)), 0.0, float(rng.uniform(-40, 40))))) + [scenes.camera_c2(), scenes.orbit_camera(int(rng.randint(50)))]
    ctx = capi.Cull(0)
    ctx.set_option(capi.OPT_KERNEL, kernel)
    ctx.set_option(capi.OPT_FUSE_LIST, fuse)
    ctx.set_option(capi.OPT_LIST_OFFSETS, offsets)
    ctx.set_option(capi.OPT_LINE_WORDS, line_words)
    ctx.set_matrices(flat)
    cur = n
    lo, ex, ti = lower4[:cur].copy(), extent4[:cur].copy(), tidx[:cur].copy()
    ctx.set_objects(lo, ex, ti)
    res = [ctx.result_create() for _ in range(nv)]
    bufs = []
    if mirror:
        for r in res:
            hb = [capi.HostBuffer(((cap + 31) // 32) * 4), capi.HostBuffer(cap * 4), capi.HostBuffer(4)]
            bufs.append(hb)
            r.set_host_mirror(hb[0].array(np.uint32), hb[1].array(np.uint32), hb[2].array(np.uint32))
    state = [port.result_resize(np.zeros(0, np.uint32), 0, cur) for _ in range(nv)]
    state_n = cur
    checked = 0
    for frame in range(int(rng.randint(2, 6))):
        what = rng.randint(4)
        if frame and what == 0:                               # grow / shrink the group (ResultBitSet::updateChanged resize rule)
            new = int(np.clip(cur + rng.randint(-cur // 2 - 1, 4000), 1, cap))
            ctx.set_object_count(new)
            if new > cur:
                idx = np.arange(cur, new, dtype=np.uint32)
                ctx.update_objects(idx, lower4[cur:new], extent4[cur:new], tidx[cur:new])
                lo = np.concatenate([lo, lower4[cur:new]]); ex = np.concatenate([ex, extent4[cur:new]]); ti = np.concatenate([ti, tidx[cur:new]])
            else:
                lo, ex, ti = lo[:new], ex[:new], ti[:new]
            cur = new
        elif frame and what == 1 and cur > 4:                 # live edits of random objects
            k = int(rng.randint(1, min(cur, 3000)))
            idx = rng.choice(cur, size=k, replace=False).astype(np.uint32)
            src = rng.randint(0, cap, size=k)
            ctx.update_objects(idx, lower4[src], extent4[src], tidx[src])
            lo[idx], ex[idx], ti[idx] = lower4[src], extent4[src], tidx[src]
        order = rng.permutation(len(cams))[:nv]
        vps = np.ascontiguousarray(np.stack([cams[j] for j in order]), np.float32)
        ctx.run(res, vps)
        for v in range(nv):
            if mirror:
                res[v].synchronize()
            want = port.cull_bits(np.ascontiguousarray(lo), np.ascontiguousarray(ex), np.ascontiguousarray(ti), flat, vps[v], threads=4)
            if cur != state_n:                                # ResultBitSet::updateChanged, the size-changed branch
                state[v] = port.result_resize(state[v], state_n, cur)
            want_changed = port.update_changed(want, state[v], cur)
            got = res[v].bits()
            assert np.array_equal(got, want), "frame %d view %d: %d bits differ" % (frame, v, int(np.unpackbits((got ^ want).view(np.uint8)).sum()))
            assert np.array_equal(res[v].changed(), want_changed), "frame %d view %d: changed list differs" % (frame, v)
            if mirror:
                assert np.array_equal(bufs[v][0].array(np.uint32)[:len(want)], want), "frame %d view %d: mirror bits" % (frame, v)
                cnt = int(bufs[v][2].array(np.uint32)[0])
                assert cnt == len(want_changed) and np.array_equal(bufs[v][1].array(np.uint32)[:cnt], want_changed), "frame %d view %d: mirror list" % (frame, v)
            checked += cur
        state_n = cur
    for r in res:
        if mirror:
            r.set_host_mirror(None, None, None)
        r.close()
    ctx.close()
    for hb in bufs:
        for b in hb:
            b.close()
    return checked


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--rounds", type=int, default=0)
    ap.add_argument("--verbose", type=int, default=0)
    a = ap.parse_args()
    port = Port()
    master = np.random.RandomState(a.seed if a.seed is not None else int(time.time()) & 0x7FFFFFFF)
    t0 = time.time()
    rounds = decisions = 0
    while (a.rounds and rounds < a.rounds) or (not a.rounds and time.time() - t0 < a.seconds):
        seed = int(master.randint(1, 1 << 30))
        lines = []
        try:
            decisions += one_round(port, np.random.RandomState(seed), lines.append)
        except Exception:
            print("FAILED in round seed %d: %s" % (seed, "; ".join(lines)))
            raise
        if a.verbose:
            print(seed, "; ".join(lines))
        rounds += 1
    print("fuzz ok: %d rounds, %d object-view decisions and their changed lists equal to the oracle (%.0f s)" % (rounds, decisions, time.time() - t0))


if __name__ == "__main__":
    main()
