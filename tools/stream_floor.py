#!/usr/bin/env python
"""Developer probe: how long does a plain read of N bytes take on this GPU when the L2 is cold - the floor a small cull
(BASELINE C2: 1 Mi objects = 100.7 MB) is measured against in DESIGN.md.  Same flush protocol as bench.py's small
workloads (256 MiB written, 256 MiB read, then the timed launch bracketed by its own event pair)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeline_b200 import capi  # noqa: E402


def main():
    s = capi.Stream()
    flush, sweep = capi.Buffer(256 << 20), capi.Buffer(256 << 20)
    sweep.fill(1)
    for mib_objects in (1, 2, 4, 16, 64):
        nbytes = mib_objects * (1 << 20) * 96
        buf = capi.Buffer(nbytes)
        buf.fill(3)
        evs = [(capi.Event(), capi.Event()) for _ in range(30)]
        for f, (a, b) in enumerate(evs):
            flush.fill(f & 0xFF, s)
            capi.read_sweep(sweep.ptr, 256 << 20, s)
            a.record(s)
            capi.read_sweep(buf.ptr, nbytes, s)
            b.record(s)
        s.sync()
        t = np.array([a.elapsed_ms(b) for a, b in evs])
        print("plain read of %6.1f MB (%2d Mi objects x 96 B), cold L2: median %.4f ms (min %.4f) = %.0f GB/s" % (
            nbytes / 1e6, mib_objects, np.median(t), t.min(), nbytes / np.median(t) / 1e6))
        buf.close()


if __name__ == "__main__":
    main()
