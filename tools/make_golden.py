#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the compiled, UNMODIFIED reference
(oracle/_ref/libdpref.so, built by `make -C oracle ref` from /root/reference) on the
deterministic inputs of tests/cases.py.

Run in the build container only (it needs /root/reference to build the library):

    python tools/make_golden.py

The reference's own tests hold no vectors for this path, so these
fixtures - answers of the reference itself - are what pins the oracle restatement and the
CUDA path on machines where /root/reference does not exist.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.loader import Port, Reference, build_ref  # noqa: E402
from pipeline_b200 import scenes  # noqa: E402
from tests import cases  # noqa: E402
from tests.engines import RefEngine, run_lifecycle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


def main():
    build_ref()
    ref = Reference()
    port = Port()
    os.makedirs(OUT, exist_ok=True)

    # 1. SURVEY.md 8(c) known-answer grid ---------------------------------------------------
    lower4, extent4, upper4, mats, tidx, vp = scenes.grid_scene(32)
    e = RefEngine(ref)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    bits, changed = e.cull(vp)
    bits2, changed2 = e.cull(vp)
    bbox = e.bounding_box()
    visible = int(np.unpackbits(bits.view(np.uint8)).sum())
    assert visible == 3100 and len(changed) == 29668 and len(changed2) == 0, (visible, len(changed))
    np.savez_compressed(os.path.join(OUT, "grid32.npz"), bits=bits, changed=changed, bbox=bbox,
                        visible=visible, fnv=np.uint64(port.fnv(bits, len(tidx))),
                        inputs=digest(lower4, upper4, mats, tidx, vp))
    e.close()

    # 2. random scene, moving camera: bits + changed list per frame --------------------------
    lower4, extent4, upper4, mats, tidx = cases.random_case()
    e = RefEngine(ref)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    out = {}
    for k, vp in enumerate(cases.frames()):
        b, c = e.cull(vp)
        out["bits%d" % k], out["changed%d" % k] = b, c
    out["bbox"] = e.bounding_box()
    out["inputs"] = digest(lower4, upper4, mats, tidx, *cases.frames())
    np.savez_compressed(os.path.join(OUT, "random20k.npz"), **out)
    e.close()

    # 2b. affine objects sitting on the clip planes, eight views (the multi-view filter's adversarial scene) -----
    n = 40000 + 37
    lower4, extent4, mats, tidx = cases.affine_boundary_scene(n, 101)
    upper4 = lower4.copy()
    upper4[:, :3] = lower4[:, :3] + extent4[:, :3]
    vps = cases.boundary_views()
    e = RefEngine(ref)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    out = {}
    for v in range(len(vps)):
        b, _ = e.cull(vps[v])
        out["bits%d" % v] = b
    out["inputs"] = digest(lower4, upper4, mats, tidx, vps)
    np.savez_compressed(os.path.join(OUT, "boundary40k.npz"), **out)
    e.close()

    # 2c. the same kind of scene under view-projections scaled from 1e-44 to 1e30 (underflow / overflow of the products)
    n = 20000 + 13
    lower4, extent4, mats, tidx = cases.affine_boundary_scene(n, 104)
    upper4 = lower4.copy()
    upper4[:, :3] = lower4[:, :3] + extent4[:, :3]
    vps = cases.scaled_views()
    e = RefEngine(ref)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    out = {}
    for v in range(len(vps)):
        b, _ = e.cull(vps[v])
        out["bits%d" % v] = b
    out["inputs"] = digest(lower4, upper4, mats, tidx, vps)
    np.savez_compressed(os.path.join(OUT, "scaledvp20k.npz"), **out)
    e.close()

    # 3. boundary / non-finite values -----------------------------------------------------------
    lower4, extent4, upper4, mats, tidx, vps = cases.special_case()
    e = RefEngine(ref)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    out = {}
    for k, vp in enumerate(vps):
        b, c = e.cull(vp)
        out["bits%d" % k], out["changed%d" % k] = b, c
    out["inputs"] = digest(lower4, upper4, mats, tidx, *vps)
    np.savez_compressed(os.path.join(OUT, "special4k.npz"), **out)
    e.close()

    # 4. gather through transformIndex with an 80-byte stride ------------------------------------
    lower4, extent4, upper4, raw, tidx, stride = cases.gather_case()
    e = RefEngine(ref)
    e.add(lower4, upper4, tidx)
    e.set_matrices(raw.reshape(-1), stride, len(raw))
    b, c = e.cull(scenes.camera_c2())
    np.savez_compressed(os.path.join(OUT, "gather5k.npz"), bits=b, changed=c, bbox=e.bounding_box(),
                        inputs=digest(lower4, upper4, raw, tidx))
    e.close()

    # 5. object lifecycle: add / remove / grow between culls -------------------------------------
    lower4, extent4, upper4, mats, tidx = cases.random_case(5000, seed=0x11FE)
    e = RefEngine(ref)
    res = run_lifecycle(e, cases.lifecycle_script(), lower4, upper4, mats, cases.frames())
    out = {"inputs": digest(lower4, upper4, mats)}
    for k, (b, c, n) in enumerate(res):
        out["bits%d" % k], out["changed%d" % k], out["count%d" % k] = b, c, n
    np.savez_compressed(os.path.join(OUT, "lifecycle.npz"), **out)
    e.close()

    # 6. transform tree: three compute() calls with different dirty sets --------------------------
    entries, offsets, n_nodes, local = cases.tree_case()
    t = ref.tree()
    order = t.add_many(entries[:, 0], local[entries[:, 1].astype(np.int64)])
    assert np.array_equal(order, entries[:, 1]), "Tree::addTransform index order differs from topology"
    out = {"inputs": digest(entries, offsets, local), "capacity": t.count()}
    t.compute()
    out["world0"], out["dirty0"] = t.world()[:n_nodes], t.dirty_world()[: (n_nodes + 31) // 32]
    for frame in (1, 2, 3):
        idx, m = cases.tree_updates(n_nodes, frame)
        t.update_locals(idx, m)
        t.compute()
        out["world%d" % frame], out["dirty%d" % frame] = t.world()[:n_nodes], t.dirty_world()[: (n_nodes + 31) // 32]
    np.savez_compressed(os.path.join(OUT, "tree.npz"), **out)
    t.close()

    for f in sorted(os.listdir(OUT)):
        print("%-16s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
