"""Deterministic synthetic scenes for the culling path.

One counter-based generator feeds the oracle, the CUDA path and the on-device generator
(``dpcuSceneGenerate`` in csrc/scene_gen.cu) with identical bytes:

    splitmix64 output k of seed s :  z = mix(s + (k+1)*0x9E3779B97F4A7C15)
    uniform float32              :  u = (z >> 40) * 2**-24          (exact in binary32)

Object ``i`` consumes draws ``16*i .. 16*i+12``:

    0..2   half size   h = 0.5 + 4.5*u          box = [-h, h]  (extent = 2h exactly)
    3..6   quaternion  q = 2*u - 1, normalised with one sqrt and four divisions
    7..9   scale       s = 0.5 + 1.5*u
    10..12 translation t = -1000 + 2000*u

    M = S * R * T  in the reference's row-vector convention (dp/math/Matmnt.h:1371-1379):
    rows 0..2 = s_r * R[r], row 3 = (t, 1).

All arithmetic is binary32 with one rounding per operation in the order written below; the
CUDA generator mirrors it operation for operation (compiled with -fmad=false), so a host
replay of any slice matches the device bytes (tests/test_scene_gen.py).

Camera helpers restate the *form* of the reference's makeFrustum / makePerspective /
makeLookAt (dp/math/Matmnt.h:1238-1369); view-projection matrices are inputs to the path, so
only determinism matters for them, not bit-equality with the reference's helpers.
"""
from __future__ import annotations

import math

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
DRAWS_PER_OBJECT = 16

SEED_C2 = 0x5EED0002
SEED_C3 = 0x5EED0003
SEED_C4 = 0x5EED0004
SEED_C5 = 0x5EED0005

f32 = np.float32


def splitmix64_at(seed: int, k: np.ndarray) -> np.ndarray:
    """k-th output (k = 0, 1, ...) of splitmix64 seeded with ``seed``; vectorised over k (uint64)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (k.astype(np.uint64) + np.uint64(1)) * GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform_at(seed: int, k: np.ndarray) -> np.ndarray:
    z = splitmix64_at(seed, k)
    return (z >> np.uint64(40)).astype(np.float32) * f32(2.0 ** -24)


def random_objects(seed: int, first: int, count: int):
    """Objects ``first .. first+count`` of the C2/C4/C5 family.

    Returns ``lower4`` (count,4: xyz, w=1), ``extent4`` (count,4: xyz, w=0), ``upper4`` and
    ``mats`` (count,4,4), all float32, and ``tidx`` = global object index (uint32).
    """
    i = np.arange(first, first + count, dtype=np.uint64)
    base = i * np.uint64(DRAWS_PER_OBJECT)

    def u(j):
        return uniform_at(seed, base + np.uint64(j))

    h = [f32(0.5) + f32(4.5) * u(j) for j in range(3)]
    lower4 = np.zeros((count, 4), dtype=np.float32)
    upper4 = np.zeros((count, 4), dtype=np.float32)
    for a in range(3):
        lower4[:, a] = -h[a]
        upper4[:, a] = h[a]
    lower4[:, 3] = 1.0
    upper4[:, 3] = 1.0
    extent4 = np.zeros((count, 4), dtype=np.float32)
    extent4[:, :3] = upper4[:, :3] - lower4[:, :3]

    qx, qy, qz, qw = [f32(2.0) * u(3 + j) - f32(1.0) for j in range(4)]
    ln = np.sqrt(((qx * qx + qy * qy) + qz * qz) + qw * qw)
    zero = ln == f32(0.0)
    ln = np.where(zero, f32(1.0), ln)
    qx, qy, qz, qw = qx / ln, qy / ln, qz / ln, qw / ln
    qw = np.where(zero, f32(1.0), qw)

    s = [f32(0.5) + f32(1.5) * u(7 + j) for j in range(3)]
    t = [f32(-1000.0) + f32(2000.0) * u(10 + j) for j in range(3)]

    two, one = f32(2.0), f32(1.0)
    xx, yy, zz = qx * qx, qy * qy, qz * qz
    xy, xz, yz = qx * qy, qx * qz, qy * qz
    wx, wy, wz = qw * qx, qw * qy, qw * qz
    R = [
        [one - two * (yy + zz), two * (xy + wz), two * (xz - wy)],
        [two * (xy - wz), one - two * (xx + zz), two * (yz + wx)],
        [two * (xz + wy), two * (yz - wx), one - two * (xx + yy)],
    ]
    mats = np.zeros((count, 4, 4), dtype=np.float32)
    for r in range(3):
        for c in range(3):
            mats[:, r, c] = s[r] * R[r][c]
    for c in range(3):
        mats[:, 3, c] = t[c]
    mats[:, 3, 3] = 1.0
    tidx = np.arange(first, first + count, dtype=np.uint32)
    return lower4, extent4, upper4, mats, tidx


# ------------------------------------------------------------------------------ cameras
def make_frustum(left, right, bottom, top, znear, zfar) -> np.ndarray:
    """Form of dp/math/Matmnt.h:1324-1344, evaluated in binary32."""
    l, r, b, t, n, f = (f32(x) for x in (left, right, bottom, top, znear, zfar))
    v0 = (r + l) / (r - l)
    v1 = (t + b) / (t - b)
    v2 = -(f + n) / (f - n)
    v3 = f32(-2.0) * f * n / (f - n)
    v4 = f32(2.0) * n / (r - l)
    v5 = f32(2.0) * n / (t - b)
    return np.array([[v4, 0, 0, 0], [0, v5, 0, 0], [v0, v1, v2, -1], [0, 0, v3, 0]], dtype=np.float32)


def make_perspective(fovy_deg, aspect, znear, zfar) -> np.ndarray:
    """Form of dp/math/Matmnt.h:1357-1369."""
    tanfov = f32(math.tan(math.radians(fovy_deg) * 0.5))
    r = tanfov * f32(aspect) * f32(znear)
    t = tanfov * f32(znear)
    return make_frustum(-r, r, -t, t, znear, zfar)


def make_look_at(eye, center, up) -> np.ndarray:
    """Form of dp/math/Matmnt.h:1250-1281 (translation premultiplied)."""
    eye = np.asarray(eye, np.float32)
    f = np.asarray(center, np.float32) - eye
    f = f / np.sqrt(np.dot(f, f)).astype(np.float32)
    s = np.cross(f, np.asarray(up, np.float32)).astype(np.float32)
    s = s / np.sqrt(np.dot(s, s)).astype(np.float32)
    u = np.cross(s, f).astype(np.float32)
    trans = np.eye(4, dtype=np.float32)
    trans[3, :3] = -eye
    ori = np.eye(4, dtype=np.float32)
    ori[:3, 0] = s
    ori[:3, 1] = u
    ori[:3, 2] = -f
    return mat_mul(trans, ori)


def mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Row-major 4x4 product with the reference's association order (Matmnt.h:1381-1415)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    r = np.zeros((4, 4), dtype=np.float32)
    for i in range(4):
        for j in range(4):
            acc = a[i, 0] * b[0, j]
            acc = acc + a[i, 1] * b[1, j]
            acc = acc + a[i, 2] * b[2, j]
            acc = acc + a[i, 3] * b[3, j]
            r[i, j] = acc
    return r


def camera_c2() -> np.ndarray:
    """C2 camera: at the origin looking down -z, 60 degree fovy, 16:9, near 1, far 1500."""
    return make_perspective(60.0, 16.0 / 9.0, 1.0, 1500.0)


def orbit_camera(frame: int, radius: float = 400.0) -> np.ndarray:
    """A moving camera for multi-frame changed-list tests and the e2e bench."""
    a = 0.05 * frame
    eye = (radius * math.sin(a), 40.0 * math.sin(0.3 * a), radius * math.cos(a))
    view = make_look_at(eye, (0.0, 0.0, 0.0), (0.0, 1.0, 0.0))
    return mat_mul(view, make_perspective(60.0, 16.0 / 9.0, 1.0, 1500.0))


def cube_map_cameras(eye=(0.0, 0.0, 0.0)) -> np.ndarray:
    """C4: six 90-degree cube-map faces from one eye, shape (6,4,4)."""
    proj = make_perspective(90.0, 1.0, 1.0, 1500.0)
    e = np.asarray(eye, np.float32)
    dirs = [((1, 0, 0), (0, -1, 0)), ((-1, 0, 0), (0, -1, 0)), ((0, 1, 0), (0, 0, 1)),
            ((0, -1, 0), (0, 0, -1)), ((0, 0, 1), (0, -1, 0)), ((0, 0, -1), (0, -1, 0))]
    out = np.zeros((6, 4, 4), dtype=np.float32)
    for k, (d, up) in enumerate(dirs):
        out[k] = mat_mul(make_look_at(e, e + np.asarray(d, np.float32), up), proj)
    return out


# ------------------------------------------------------------------------------ C1 grid (KAT)
def grid_scene(g: int = 32):
    """SURVEY.md 8(c) known-answer scene: g^3 unit cubes, one matrix per object."""
    n = g ** 3
    idx = np.arange(n)
    x, y, z = idx % g, (idx // g) % g, idx // (g * g)
    lower4 = np.tile(np.array([-0.5, -0.5, -0.5, 1.0], dtype=np.float32), (n, 1))
    upper4 = np.tile(np.array([0.5, 0.5, 0.5, 1.0], dtype=np.float32), (n, 1))
    extent4 = np.zeros((n, 4), dtype=np.float32)
    extent4[:, :3] = upper4[:, :3] - lower4[:, :3]
    mats = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    mats[:, 3, 0] = 2.0 * x - g
    mats[:, 3, 1] = 2.0 * y - g
    mats[:, 3, 2] = 2.0 * z - g
    tidx = np.arange(n, dtype=np.uint32)
    view = np.eye(4, dtype=np.float32)
    view[3, 2] = -20.0
    vp = mat_mul(view, make_frustum(-0.5, 0.5, -0.5, 0.5, 1.0, 40.0))
    return lower4, extent4, upper4, mats, tidx, vp


# ------------------------------------------------------------------------------ C3 hierarchy
def hierarchy_topology(levels=(4096, 65536, 1048576, 16777216)):
    """Level-sorted {parent, transform} entries for a uniform-fan-out tree under the virtual
    root 0 (dp/transform/src/Tree.cpp:42-47).  Node indices are assigned level by level from 1,
    which is what Tree::addTransform produces when nodes are added in that order.

    Returns entries (E,2) uint32, level_offsets (L+1,) uint32, n_nodes (incl. root).
    """
    entries = []
    offsets = [0]
    first_prev, n_prev = 0, 1
    nxt = 1
    for n in levels:
        t = np.arange(nxt, nxt + n, dtype=np.uint32)
        fan = n // n_prev
        parent = (first_prev + (np.arange(n, dtype=np.uint64) // np.uint64(fan))).astype(np.uint32)
        entries.append(np.stack([parent, t], axis=1))
        offsets.append(offsets[-1] + n)
        first_prev, n_prev = nxt, n
        nxt += n
    return np.concatenate(entries).astype(np.uint32), np.asarray(offsets, dtype=np.uint32), nxt


def hierarchy_locals(seed: int, first: int, count: int, frame: int = 0) -> np.ndarray:
    """Local matrices of nodes first..first+count for C3: rotation about z by a small
    per-node angle that advances with the frame index, plus a per-node translation."""
    i = np.arange(first, first + count, dtype=np.uint64)
    base = i * np.uint64(4)
    ang = (uniform_at(seed, base) - f32(0.5)) * f32(0.2) + f32(0.01) * f32(frame)
    c, s = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)
    m = np.zeros((count, 4, 4), dtype=np.float32)
    m[:, 0, 0], m[:, 0, 1] = c, s
    m[:, 1, 0], m[:, 1, 1] = -s, c
    m[:, 2, 2] = 1.0
    m[:, 3, 3] = 1.0
    for a in range(3):
        m[:, 3, a] = (uniform_at(seed, base + np.uint64(1 + a)) - f32(0.5)) * f32(40.0)
    return m
