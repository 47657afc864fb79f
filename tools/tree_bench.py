#!/usr/bin/env python
"""Developer probe: K1 (transform propagation) alone on the C3 tree, both kernel forms."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeline_b200 import capi, scenes  # noqa: E402


def main():
    levels = (4096, 65536, 1048576, 16777216)
    entries, offsets, n_nodes = scenes.hierarchy_topology(levels)
    for wide_min in (1 << 16, 1 << 40, 0):
        tree = capi.Tree(0)
        tree.set_option(capi.TREE_OPT_WIDE_MIN_NODES, wide_min)
        tree.set_topology(entries, offsets, n_nodes)
        lptr, _ = tree.local_ptr()
        lo, ex = capi.Buffer(n_nodes * 16), capi.Buffer(n_nodes * 16)
        capi.scene_generate(scenes.SEED_C3 + 1, 0, n_nodes, 0, lo.ptr, ex.ptr, lptr)
        capi.device_sync()
        lo.close(), ex.close()
        s = capi.Stream()
        e0, e1 = capi.Event(), capi.Event()
        times = []
        for it in range(8):
            tree.mark_dirty(1, n_nodes - 1)
            capi.device_sync()
            e0.record(s)
            tree.compute(s)
            e1.record(s)
            s.sync()
            times.append(e0.elapsed_ms(e1))
        t = float(np.median(times[2:]))
        alg = sum(levels) * 136.0 + (1 + sum(levels[:-1])) * 64.0
        crc = int(np.bitwise_xor.reduce(tree.world(n_nodes - 4096, 4096).view(np.uint32).reshape(-1)))
        print("wide_min=%d: Tree::compute over %d nodes all dirty: median %.4f ms -> %.0f GB/s algorithmic (%.2f GB); xor %08x"
              % (wide_min, n_nodes - 1, t, alg / t / 1e6, alg / 1e9, crc))
        tree.close()


if __name__ == "__main__":
    main()
