#!/usr/bin/env python
"""Developer probe: how much slack does the multi-view filter's margin have?  64 Mi objects x 6 views with the
filter off (reference arithmetic for every pair), on, with 1/8 of the margin and with no margin at all."""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeline_b200 import capi, scenes  # noqa: E402

n = 1 << 26
lo, ex, mt = capi.Buffer(n * 16), capi.Buffer(n * 16), capi.Buffer(n * 64)
capi.scene_generate(scenes.SEED_C4, 0, n, 0, lo.ptr, ex.ptr, mt.ptr)
capi.device_sync()
ctx = capi.Cull(0)
ctx.set_objects(lo.ptr, ex.ptr, None, capi.MEM_DEVICE, n=n)
ctx.bind_matrices(mt.ptr, n)
ref = {}
for eye in ((0.0, 0.0, 0.0), (3.0, -7.0, 11.0), (250.0, 100.0, -40.0)):
    vps = np.ascontiguousarray(scenes.cube_map_cameras(eye), np.float32)
    for mode in (0, 1, 2, 3):
        ctx.set_option(capi.OPT_FILTER, mode)
        res = [ctx.result_create() for _ in range(6)]
        ctx.run(res, vps)
        bits = [r.bits() for r in res]
        for r in res:
            r.close()
        if mode == 0:
            ref[eye] = bits
            continue
        diff = sum(int(np.unpackbits((a ^ b).view(np.uint8)).sum()) for a, b in zip(bits, ref[eye]))
        print("eye %s filter mode %d (%s): %d of %d decisions differ from the reference arithmetic"
              % (eye, mode, {1: "margin 2^-17 S", 2: "margin 2^-20 S", 3: "no margin"}[mode], diff, 6 * n))
