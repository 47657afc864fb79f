"""CPU tests: the C restatement (oracle/ref_cull.c) is pinned against
 (1) golden vectors produced by the compiled unmodified reference (tests/golden, tools/make_golden.py)
 (2) the compiled reference itself, live, when oracle/_ref/libdpref.so is present."""
import hashlib

import numpy as np
import pytest

from pipeline_b200 import scenes
from tests import cases
from tests.engines import PortEngine, RefEngine, run_lifecycle


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


def popcount(words):
    return int(np.unpackbits(words.view(np.uint8)).sum())


# ------------------------------------------------------------------ golden vectors
def test_grid_known_answer(port, golden):
    """32^3 unit cubes -> 3100 visible, 29668 changed on the first cull."""
    g = golden("grid32")
    lower4, extent4, upper4, mats, tidx, vp = scenes.grid_scene(32)
    assert np.array_equal(digest(lower4, upper4, mats, tidx, vp), g["inputs"]), "scene generator drifted"
    e = PortEngine(port)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    bits, changed = e.cull(vp)
    assert popcount(bits) == 3100 == int(g["visible"])
    assert len(changed) == 29668
    assert np.array_equal(bits, g["bits"])
    assert np.array_equal(changed, g["changed"])
    assert port.fnv(bits, len(tidx)) == int(g["fnv"])
    bits2, changed2 = e.cull(vp)
    assert len(changed2) == 0 and np.array_equal(bits2, bits)
    assert np.array_equal(e.bounding_box(), g["bbox"])
    assert np.array_equal(g["bbox"], np.array([-32.5, -32.5, -32.5, 30.5, 30.5, 30.5], np.float32))


def test_random_scene_moving_camera(port, golden):
    g = golden("random20k")
    lower4, extent4, upper4, mats, tidx = cases.random_case()
    assert np.array_equal(digest(lower4, upper4, mats, tidx, *cases.frames()), g["inputs"])
    e = PortEngine(port)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    for k, vp in enumerate(cases.frames()):
        bits, changed = e.cull(vp)
        assert np.array_equal(bits, g["bits%d" % k]), "frame %d bits" % k
        assert np.array_equal(changed, g["changed%d" % k]), "frame %d changed list" % k
        assert 0 < popcount(bits) < len(tidx)
    assert np.array_equal(e.bounding_box(), g["bbox"])


def test_special_values(port, golden):
    """Corners exactly on planes, NaN / Inf / denormal matrix entries, zero extents, w <= 0."""
    g = golden("special4k")
    lower4, extent4, upper4, mats, tidx, vps = cases.special_case()
    assert np.array_equal(digest(lower4, upper4, mats, tidx, *vps), g["inputs"])
    e = PortEngine(port)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    for k, vp in enumerate(vps):
        bits, changed = e.cull(vp)
        assert np.array_equal(bits, g["bits%d" % k])
        assert np.array_equal(changed, g["changed%d" % k])


def test_boundary_scene_eight_views(port, golden):
    """the multi-view filter's adversarial scene (objects sitting on the clip planes): the restatement equals the
    compiled reference's answers for all eight views"""
    g = golden("boundary40k")
    lower4, extent4, mats, tidx = cases.affine_boundary_scene(40037, 101)
    vps = cases.boundary_views()
    for v in range(8):
        assert np.array_equal(port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), vps[v]), g["bits%d" % v]), v


def test_scaled_view_projections(port, golden):
    """view-projections scaled from 1e-44 to 1e30 (products underflowing to denormals / heading for overflow): the
    restatement equals the compiled reference's answers for all 56 views"""
    g = golden("scaledvp20k")
    lower4, extent4, mats, tidx = cases.affine_boundary_scene(20013, 104)
    vps = cases.scaled_views()
    assert len(vps) == 4 * len(cases.VP_SCALES)
    differing = 0
    for v in range(len(vps)):
        want = g["bits%d" % v]
        assert np.array_equal(port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), vps[v]), want), v
        differing += int(not np.array_equal(want, g["bits%d" % (v - v % len(cases.VP_SCALES) + 8)]))
    assert differing > 0          # the scaling really changes answers somewhere (else this golden pins nothing new)


def test_gather_and_stride(port, golden):
    g = golden("gather5k")
    lower4, extent4, upper4, raw, tidx, stride = cases.gather_case()
    assert np.array_equal(digest(lower4, upper4, raw, tidx), g["inputs"])
    e = PortEngine(port)
    e.add(lower4, upper4, tidx)
    e.set_matrices(raw.reshape(-1), stride, len(raw))
    bits, changed = e.cull(scenes.camera_c2())
    assert np.array_equal(bits, g["bits"])
    assert np.array_equal(changed, g["changed"])
    assert np.array_equal(e.bounding_box(), g["bbox"])


def test_lifecycle(port, golden):
    g = golden("lifecycle")
    lower4, extent4, upper4, mats, tidx = cases.random_case(5000, seed=0x11FE)
    assert np.array_equal(digest(lower4, upper4, mats), g["inputs"])
    res = run_lifecycle(PortEngine(port), cases.lifecycle_script(), lower4, upper4, mats, cases.frames())
    for k, (bits, changed, n) in enumerate(res):
        assert n == int(g["count%d" % k])
        assert np.array_equal(bits, g["bits%d" % k]), "step %d bits" % k
        assert np.array_equal(changed, g["changed%d" % k]), "step %d changed" % k


def test_tree_golden(port, golden):
    g = golden("tree")
    entries, offsets, n_nodes, local = cases.tree_case()
    assert np.array_equal(digest(entries, offsets, local), g["inputs"])
    nw = (n_nodes + 31) // 32
    world = np.zeros_like(local)
    world[0] = np.eye(4, dtype=np.float32)
    dirty_local = np.zeros(nw, np.uint32)
    dirty_world = np.zeros(nw, np.uint32)
    # Tree() marks the root's local dirty (Tree.cpp:42), addTransform marks every new node (Tree.cpp:63)
    for i in range(n_nodes):
        dirty_local[i >> 5] |= np.uint32(1 << (i & 31))
    port.tree_compute(local, world, entries, offsets, dirty_local, dirty_world)
    assert np.array_equal(world, g["world0"])
    assert np.array_equal(dirty_world, g["dirty0"])
    for frame in (1, 2, 3):
        dirty_world[:] = 0
        idx, m = cases.tree_updates(n_nodes, frame)
        local[idx.astype(np.int64)] = m
        for i in idx:
            dirty_local[int(i) >> 5] |= np.uint32(1 << (int(i) & 31))
        port.tree_compute(local, world, entries, offsets, dirty_local, dirty_world)
        assert np.array_equal(world, g["world%d" % frame]), "frame %d world" % frame
        assert np.array_equal(dirty_world, g["dirty%d" % frame]), "frame %d dirty set" % frame
        assert not dirty_local.any()


def test_tree_known_answer(reference):
    """root -> A(M_5) -> B(M_7): world[B] row 3 = (-40,-64,-64,1), capacity 65536."""
    def M(i):
        m = np.eye(4, dtype=np.float32)
        m[3, :3] = [2 * (i % 32) - 32, 2 * ((i // 32) % 32) - 32, 2 * (i // 1024) - 32]
        return m
    t = reference.tree()
    a = t.add(0, M(5))
    b = t.add(a, M(7))
    t.compute()
    assert t.count() == 65536
    assert np.array_equal(t.world()[b][3], np.array([-40, -64, -64, 1], np.float32))


# ------------------------------------------------------------------ live reference
def test_port_matches_live_reference_random(port, reference):
    """1 Mi random objects (BASELINE config C2) + boundary scene, port vs compiled reference."""
    n = 1 << 20
    lower4, extent4, upper4, mats, tidx = scenes.random_objects(scenes.SEED_C2, 0, n)
    r = RefEngine(reference)
    p = PortEngine(port)
    for e in (r, p):
        e.add(lower4, upper4, tidx)
        e.set_matrices(mats.reshape(-1))
    for vp in cases.frames(3):
        rb, rc = r.cull(vp)
        pb, pc = p.cull(vp)
        assert np.array_equal(rb, pb)
        assert np.array_equal(rc, pc)
    vis = popcount(rb) / n
    assert 0.02 < vis < 0.5, vis
    assert np.array_equal(r.bounding_box(), p.bounding_box())
    r.close()


def test_port_matches_live_reference_fuzz(port, reference):
    """Random special-value scenes with several seeds."""
    for seed in range(20, 28):
        lower4, extent4, upper4, mats, tidx, vps = cases.special_case(2048, seed=seed)
        r = RefEngine(reference)
        p = PortEngine(port)
        for e in (r, p):
            e.add(lower4, upper4, tidx)
            e.set_matrices(mats.reshape(-1))
        for vp in vps + [scenes.camera_c2()]:
            rb, rc = r.cull(vp)
            pb, pc = p.cull(vp)
            assert np.array_equal(rb, pb), seed
            assert np.array_equal(rc, pc), seed
        r.close()


def test_port_multithreaded_equals_single(port):
    lower4, extent4, upper4, mats, tidx = cases.random_case(100003)
    a = port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), scenes.camera_c2())
    b = port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), scenes.camera_c2(), threads=5)
    assert np.array_equal(a, b)


def test_matrix_dirty_protocol_live_reference(port, reference):
    """Changing matrices in place needs groupMatrixChanged on the reference (SURVEY hard part 5)."""
    lower4, extent4, upper4, mats, tidx = cases.random_case(4096)
    mats = mats.copy()
    r = RefEngine(reference)
    p = PortEngine(port)
    for e in (r, p):
        e.add(lower4, upper4, tidx)
        e.set_matrices(mats.reshape(-1))
    vp = scenes.camera_c2()
    r.cull(vp), p.cull(vp)
    idx = np.arange(0, 4096, 7, dtype=np.uint32)
    mats[idx.astype(np.int64), 3, 2] -= np.float32(300.0)
    for e in (r, p):
        e.matrices_changed(idx)
    rb, rc = r.cull(vp)
    pb, pc = p.cull(vp)
    assert np.array_equal(rb, pb) and np.array_equal(rc, pc) and len(rc) > 0
    r.close()
