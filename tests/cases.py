"""Deterministic test-case inputs shared by tools/make_golden.py (which runs the compiled
reference on them and stores its answers under tests/golden/) and by the parity tests."""
from __future__ import annotations

import numpy as np

from pipeline_b200 import scenes

f32 = np.float32


def random_case(n=20000, seed=scenes.SEED_C2, first=0):
    lower4, extent4, upper4, mats, tidx = scenes.random_objects(seed, first, n)
    tidx = (tidx - np.uint32(first)).astype(np.uint32)
    return lower4, extent4, upper4, mats, tidx


def frames(k=5):
    return [scenes.camera_c2()] + [scenes.orbit_camera(f) for f in range(1, k)]


def special_case(n=4096, seed=7):
    """Boundary / non-finite inputs: coordinates exactly on +-w, NaN / Inf matrix entries,
    zero extents, negative scales, objects behind the camera (w <= 0), denormals."""
    rng = np.random.RandomState(seed)
    vals = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 2.0, -2.0, 3.0, 1e-40, -1e-40, 1e30, -1e30,
                     np.inf, -np.inf, np.nan, 1.0000001, 0.99999994, 16777216.0], dtype=np.float32)
    common = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 2.0, -2.0], dtype=np.float32)
    lower4 = np.zeros((n, 4), dtype=np.float32)
    upper4 = np.zeros((n, 4), dtype=np.float32)
    lo = rng.choice(common, size=(n, 3))
    ex = np.abs(rng.choice(common, size=(n, 3)))
    lower4[:, :3] = lo
    upper4[:, :3] = lo + ex
    lower4[:, 3] = 1.0
    upper4[:, 3] = 1.0
    mats = rng.choice(common, size=(n, 4, 4)).astype(np.float32)
    # sprinkle the nasty values into a fraction of the matrices
    mask = rng.rand(n, 4, 4) < 0.08
    mats[mask] = rng.choice(vals, size=int(mask.sum()))
    # a block of plain affine matrices with integer translations so many corners land exactly on planes
    k = n // 4
    mats[:k] = np.eye(4, dtype=np.float32)
    mats[:k, 3, :3] = rng.randint(-4, 5, size=(k, 3)).astype(np.float32)
    extent4 = np.zeros((n, 4), dtype=np.float32)
    extent4[:, :3] = upper4[:, :3] - lower4[:, :3]
    tidx = np.arange(n, dtype=np.uint32)
    # simple projection: w = -z (perspective), x,y scale 1, z' = z  -> planes at |x|,|y|,|z'| = w
    vp = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, -1], [0, 0, 0, 0]], dtype=np.float32)
    vp2 = np.eye(4, dtype=np.float32)          # orthographic unit cube, w = 1
    return lower4, extent4, upper4, mats, tidx, [vp, vp2]


def gather_case(n=5000, m=37, stride=80, seed=11):
    """Many objects sharing few matrices, non-64-byte stride (GroupBitSet.cpp:121-135)."""
    lower4, extent4, upper4, mats, _ = scenes.random_objects(0xABCDEF, 0, n)
    rng = np.random.RandomState(seed)
    tidx = rng.randint(0, m, size=n).astype(np.uint32)
    raw = np.zeros((m, stride // 4), dtype=np.float32)
    raw[:, :16] = mats[:m].reshape(m, 16)
    raw[:, 16:] = np.nan                              # padding must never be read
    return lower4, extent4, upper4, raw, tidx, stride


def tree_case(levels=(4, 16, 64, 256), seed=scenes.SEED_C3):
    entries, offsets, n_nodes = scenes.hierarchy_topology(levels)
    local = np.zeros((n_nodes, 4, 4), dtype=np.float32)
    local[0] = np.eye(4, dtype=np.float32)
    local[1:] = scenes.hierarchy_locals(seed, 1, n_nodes - 1, frame=0)
    return entries, offsets, n_nodes, local


def tree_updates(n_nodes, frame, seed=scenes.SEED_C3):
    """Frame 1: every node dirty.  Frame 2: a sparse set.  Frame 3: nothing."""
    if frame == 1:
        idx = np.arange(1, n_nodes, dtype=np.uint32)
    elif frame == 2:
        rng = np.random.RandomState(5)
        idx = np.unique(rng.randint(1, n_nodes, size=max(1, n_nodes // 20))).astype(np.uint32)
    else:
        idx = np.zeros(0, dtype=np.uint32)
    mats = np.zeros((len(idx), 4, 4), dtype=np.float32)
    if len(idx):
        allm = scenes.hierarchy_locals(seed, 1, n_nodes - 1, frame=frame)
        mats = allm[idx.astype(np.int64) - 1]
    return idx, mats


def lifecycle_script():
    """(op, args) steps exercising add / cull / remove / resize of the result (a11, a12)."""
    return [
        ("add", 0, 3000),
        ("cull", 0),
        ("cull", 1),
        ("remove", [5, 17, 2999 - 2, 100, 0]),   # group indices at the time of each removal
        ("cull", 1),
        ("add", 3000, 1500),
        ("cull", 2),
        ("remove", [4400, 31, 32, 63, 64]),
        ("add", 4500, 70),
        ("cull", 2),
        ("cull", 3),
    ]


# ------------------------------------------------------------------ scenes built to sit on the clip planes
def affine_boundary_scene(n, seed):
    """Affine objects (the filter only engages on them) built to sit ON the decision boundaries: integer / half
    translations and extents under projections with small integer entries (corners exactly on x = +-w), boxes as
    large as the frustum, boxes straddling the eye, zero extents, negative and wildly different scales."""
    rng = np.random.RandomState(seed)
    half = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 8.0], np.float32)
    lower4 = np.zeros((n, 4), np.float32)
    upper4 = np.zeros((n, 4), np.float32)
    lo = -rng.choice(half, size=(n, 3))
    ex = rng.choice(half, size=(n, 3)) + rng.choice(half, size=(n, 3))
    big = rng.rand(n) < 0.1                                   # boxes of the size of the scene
    ex[big] *= 16.0
    lower4[:, :3], upper4[:, :3] = lo, lo + ex
    lower4[:, 3] = upper4[:, 3] = 1.0
    mats = np.zeros((n, 4, 4), np.float32)
    scale = rng.choice(np.array([1.0, -1.0, 0.5, 2.0, 0.0, 1e-20, 1e12, 3.0], np.float32), size=(n, 3),
                       p=[0.4, 0.1, 0.1, 0.1, 0.05, 0.05, 0.05, 0.15])
    perm = np.array([[0, 1, 2], [1, 2, 0], [2, 0, 1], [0, 2, 1]])[rng.randint(0, 4, size=n)]
    for r in range(3):
        mats[np.arange(n), r, perm[:, r]] = scale[:, r]      # signed / scaled axis permutations: exact arithmetic
    shear = rng.rand(n) < 0.3
    mats[shear, 0, 1] = rng.choice(np.array([0.5, -0.5, 1.0], np.float32), size=int(shear.sum()))
    rot = rng.rand(n) < 0.3                                   # and a share of generic rotations
    ang = rng.rand(n).astype(np.float32) * 6.28
    c, s_ = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)
    mats[rot, 0, 0], mats[rot, 0, 1], mats[rot, 1, 0], mats[rot, 1, 1] = c[rot], s_[rot], -s_[rot], c[rot]
    mats[rot, 0, 2] = mats[rot, 1, 2] = mats[rot, 2, 0] = mats[rot, 2, 1] = 0.0
    mats[rot, 2, 2] = 1.0
    mats[:, 3, :3] = rng.randint(-24, 25, size=(n, 3)).astype(np.float32) * rng.choice(np.array([0.5, 1.0, 1.0, 4.0], np.float32), size=(n, 1))
    mats[:, 3, 3] = 1.0
    extent4 = np.zeros((n, 4), np.float32)
    extent4[:, :3] = upper4[:, :3] - lower4[:, :3]
    return lower4, extent4, mats, np.arange(n, dtype=np.uint32)


def boundary_views():
    persp = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, -1], [0, 0, -2, 0]], np.float32)        # w = -z, z' = -z - 2: integer planes
    ortho = np.eye(4, dtype=np.float32)                                                               # w = 1: the unit cube
    ortho8 = np.diag(np.array([0.125, 0.125, 0.125, 1.0], np.float32))                                # the cube [-8, 8]^3
    behind = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1], [0, 0, -2, 0]], np.float32)         # the same looking down +z
    shifted = persp.copy()
    shifted[3] = [2.0, -3.0, 0.0, 16.0]                                                               # eye moved, integer entries
    frustum = scenes.mat_mul(scenes.make_look_at((0, 0, 20), (0, 0, 0), (0, 1, 0)), scenes.make_frustum(-0.5, 0.5, -0.5, 0.5, 1.0, 40.0))
    cube = scenes.cube_map_cameras((1.0, 2.0, 3.0))
    return np.ascontiguousarray(np.stack([persp, ortho, ortho8, behind, shifted, frustum, cube[0], cube[3]]), np.float32)


VP_SCALES = (1e-44, 1e-40, 1e-38, 1e-35, 1e-31, 1e-30, 1e-29, 1e-20, 1e-8, 1e8, 1e11, 1e12, 1e20, 1e30)


def scaled_views():
    """View-projections multiplied by a common factor over 75 orders of magnitude (perspective and orthographic).
    The factor does not change the exact result - every clip coordinate scales with it - but it moves the
    reference's binary32 products into the denormal range at one end (absolute, not relative, rounding error) and
    towards overflow at the other: the multi-view filter must stand down where its relative margin no longer
    covers the error (cull_filter_pairs.cuh; the host switches it off below 2^-100 and above 2^39)."""
    base = boundary_views()
    out = []
    for k in (0, 2, 5, 6):                    # integer-plane perspective, orthographic cube, look-at frustum, cube-map face
        for sc in VP_SCALES:
            out.append((base[k].astype(np.float64) * sc).astype(np.float32))
    return np.ascontiguousarray(np.stack(out), np.float32)
