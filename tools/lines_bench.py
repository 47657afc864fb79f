#!/usr/bin/env python
"""Developer probe: the line-granular kernel with the changed list built in-kernel (look-back) vs by the
compaction kernel, device-resident and with a host mirror."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeline_b200 import capi, scenes  # noqa: E402


def main():
    n = 1 << 26
    views = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    lo, ex, mt = capi.Buffer(n * 16), capi.Buffer(n * 16), capi.Buffer(n * 64)
    capi.scene_generate(scenes.SEED_C4, 0, n, 0, lo.ptr, ex.ptr, mt.ptr)
    capi.device_sync()
    if views == 1:
        cams = [np.ascontiguousarray(scenes.orbit_camera(i), np.float32).reshape(1, 16) for i in range(45)]
    else:
        cams = [np.ascontiguousarray(scenes.cube_map_cameras((3.0 * i, 0, 0))[:views], np.float32) for i in range(45)]
    for fuse in (1, 0):
        for mirror in (0, 1):
            ctx = capi.Cull(0)
            ctx.set_objects(lo.ptr, ex.ptr, None, capi.MEM_DEVICE, n=n)
            ctx.bind_matrices(mt.ptr, n)
            ctx.set_option(capi.OPT_KERNEL, capi.KERNEL_LINES)
            ctx.set_option(capi.OPT_FUSE_LIST, fuse)
            ctx.set_option(capi.OPT_PROFILE, 1)
            res = [ctx.result_create() for _ in range(views)]
            bufs = []
            if mirror:
                for r in res:
                    hb = [capi.HostBuffer(n // 8), capi.HostBuffer(n * 4), capi.HostBuffer(4)]
                    bufs.append(hb)
                    r.set_host_mirror(hb[0].array(np.uint32), hb[1].array(np.uint32), hb[2].array(np.uint32))
            s = capi.Stream()
            e0, e1 = capi.Event(), capi.Event()
            for i in range(5):
                ctx.run(res, cams[i], s)
            s.sync()
            ctx.kernel_time()
            e0.record(s)
            for i in range(5, 35):
                ctx.run(res, cams[i], s)
                if mirror:
                    for r in res:
                        r.synchronize()
            e1.record(s)
            s.sync()
            k, kn = ctx.kernel_time()
            print("views=%d fuse_list=%d mirror=%d: step %.4f ms, lines kernel %.4f ms, changed %d"
                  % (views, fuse, mirror, e0.elapsed_ms(e1) / 30, k / kn, res[0].changed_count()))
            for r in res:
                r.close()
            ctx.close()
            for hb in bufs:
                for b in hb:
                    b.close()


if __name__ == "__main__":
    main()
