#!/usr/bin/env python
"""Developer probe (not the contract bench): device-resident cull timings for a few sizes / variants."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeline_b200 import capi, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 26)
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--fma", type=int, default=0)
    ap.add_argument("--changed", type=int, default=1)
    ap.add_argument("--filter", type=int, default=1, help="multi-view filter (DPCU_CULL_OPT_FILTER)")
    ap.add_argument("--line-words", type=int, default=0, help="DPCU_CULL_OPT_LINE_WORDS")
    ap.add_argument("--l2-prefetch", type=int, default=1, help="DPCU_CULL_OPT_L2_PREFETCH")
    ap.add_argument("--list-offsets", type=int, default=0, help="DPCU_CULL_OPT_LIST_OFFSETS")
    ap.add_argument("--profile", type=int, default=1, help="0: no per-launch events (they sit between the cull and the dependent compaction launch)")
    ap.add_argument("--static", type=int, default=0, help="1: same camera every iteration")
    ap.add_argument("--eye", type=float, nargs=3, default=None, help="multi-view: cube-map eye position")
    ap.add_argument("--move", type=float, default=0.0, help="multi-view: the eye moves by this much in x per iteration")
    ap.add_argument("--flush", type=int, default=0, help="1: write 256 MiB between iterations (cold, dirty L2); 2: ... then read another 256 MiB (cold, clean L2)")
    a = ap.parse_args()
    n = a.n
    lo, ex, mt = capi.Buffer(n * 16), capi.Buffer(n * 16), capi.Buffer(n * 64)
    capi.scene_generate(scenes.SEED_C4, 0, n, 0, lo.ptr, ex.ptr, mt.ptr)
    capi.device_sync()
    ctx = capi.Cull(0)
    ctx.set_objects(lo.ptr, ex.ptr, None, capi.MEM_DEVICE, n=n)
    ctx.bind_matrices(mt.ptr, n)
    ctx.set_option(capi.OPT_KERNEL, a.kernel)
    ctx.set_option(capi.OPT_CTAS_PER_SM, a.ctas)
    ctx.set_option(capi.OPT_FMA, a.fma)
    ctx.set_option(capi.OPT_CHANGED_LIST, a.changed)
    ctx.set_option(capi.OPT_FILTER, a.filter)
    ctx.set_option(capi.OPT_LIST_OFFSETS, a.list_offsets)
    ctx.set_option(capi.OPT_LINE_WORDS, a.line_words)
    ctx.set_option(capi.OPT_L2_PREFETCH, a.l2_prefetch)
    res = [ctx.result_create() for _ in range(a.views)]
    cams = np.concatenate([scenes.cube_map_cameras(), scenes.cube_map_cameras((50.0, 20.0, -30.0))]) if a.views > 1 else None
    s = capi.Stream()
    e0, e1 = capi.Event(), capi.Event()

    def vps(f):
        if a.views > 1 and (a.eye is not None or a.move):
            e = a.eye or (0.0, 0.0, 0.0)
            return scenes.cube_map_cameras((e[0] + a.move * f, e[1], e[2]))[:a.views]
        if a.views > 1:
            return cams[:a.views]
        return scenes.camera_c2() if a.static else scenes.orbit_camera(f)

    cam = [np.ascontiguousarray(vps(f), np.float32) for f in range(3 + a.iters)]   # camera math stays outside the timed region
    for f in range(3):
        ctx.run(res, cam[f], s)
    s.sync()
    times = []
    ctx.set_option(capi.OPT_PROFILE, a.profile)
    ctx.kernel_time()
    flush = capi.Buffer(256 << 20) if a.flush else None
    sweep = capi.Buffer(256 << 20) if a.flush == 2 else None
    # everything is queued without a host synchronisation in between: with a flush (tens of microseconds of device work)
    # in front of every step the launches are enqueued ahead of the GPU, so a step's event pair brackets device time, not
    # the CPU's launch overhead
    evs = [(capi.Event(), capi.Event()) for _ in range(a.iters)]
    for f in range(a.iters):
        if flush is not None:
            flush.fill(f & 0xFF, s)
            if sweep is not None:
                capi.read_sweep(sweep.ptr, 256 << 20, s)
        evs[f][0].record(s)
        ctx.run(res, cam[3 + f], s)
        evs[f][1].record(s)
    s.sync()
    times = [x.elapsed_ms(y) for x, y in evs]
    t = np.median(times)
    k_ms, k_n = ctx.kernel_time()
    nchg = sum(r.changed_count() for r in res) if a.changed else 0
    bytes_alg = n * (96 + 0.25 * a.views) + 4 * nchg
    import zlib
    crc = 0
    for r in res:
        crc = zlib.crc32(r.bits().tobytes(), crc)
    vis = int(np.unpackbits(res[0].bits().view(np.uint8)).sum())
    print("n=%d views=%d kernel=%d ctas=%d fma=%d changed=%d: median %.4f ms (min %.4f) -> %.2f Gobj/s, %.0f GB/s algorithmic; visible %.1f%% changed %d bits-crc %08x; cull kernel avg %.4f ms"
          % (n, a.views, a.kernel, a.ctas, a.fma, a.changed, t, min(times), n / t / 1e6, bytes_alg / t / 1e6, 100.0 * vis / n, nchg, crc, k_ms / max(k_n, 1)))


if __name__ == "__main__":
    main()
