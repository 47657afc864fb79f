"""TEST INFRASTRUCTURE - ctypes loaders for the parity oracle.

Two libraries, both CPU-only and both *checkers*, never the product:

* ``libdporacle.so``   - ``oracle/ref_cull.c``, the plain-C restatement ("port").
* ``_ref/libdpref.so`` - the unmodified reference ``dp::culling::cpu::Manager`` and
  ``dp::transform::Tree`` behind ``oracle/ref_shim.cpp`` ("reference").  It is compiled in the
  build container (``make -C oracle ref``) and travels to the GPU box as a prebuilt file.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libdporacle.so")
REF_SO = os.path.join(HERE, "_ref", "libdpref.so")

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def _up(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u32p)


def build_port() -> str:
    """Compile the C restatement if it is missing or stale (gcc only, a second or two)."""
    src = os.path.join(HERE, "ref_cull.c")
    hdr = os.path.join(HERE, "ref_cull.h")
    if (not os.path.exists(PORT_SO)
            or os.path.getmtime(PORT_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return PORT_SO


def build_ref() -> str | None:
    """Compile the real reference when /root/reference is present; otherwise use the prebuilt file."""
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return REF_SO if os.path.exists(REF_SO) else None


class Port:
    """The C restatement (``kind = "port"``)."""

    def __init__(self):
        self.lib = C.CDLL(build_port())
        L = self.lib
        L.dporacle_vec4_mul_mat44.argtypes = [_f32p, _f32p, _f32p]
        L.dporacle_mat44_mul.argtypes = [_f32p, _f32p, _f32p]
        L.dporacle_box_extent.argtypes = [_f32p, _f32p, C.c_size_t, _f32p]
        L.dporacle_obb.argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.dporacle_is_visible.argtypes = [_f32p, _f32p]
        L.dporacle_is_visible.restype = C.c_int
        L.dporacle_cull_bits.argtypes = [_f32p, _f32p, _u32p, C.c_size_t, C.c_void_p, C.c_size_t, _f32p, _u32p]
        L.dporacle_cull_bits_mt.argtypes = L.dporacle_cull_bits.argtypes + [C.c_int]
        L.dporacle_result_resize.argtypes = [_u32p, C.c_size_t, C.c_size_t]
        L.dporacle_update_changed.argtypes = [_u32p, _u32p, C.c_size_t, _u32p]
        L.dporacle_update_changed.restype = C.c_size_t
        L.dporacle_result_move_bit.argtypes = [_u32p, C.c_size_t, C.c_size_t, C.c_size_t]
        L.dporacle_bounding_box.argtypes = [_f32p, _f32p, _u32p, C.c_size_t, C.c_void_p, C.c_size_t, _f32p]
        L.dporacle_tree_compute.argtypes = [_f32p, _f32p, _u32p, _u32p, C.c_int, _u32p, _u32p, C.c_size_t]
        L.dporacle_visibility_fnv1a.argtypes = [_u32p, C.c_size_t]
        L.dporacle_visibility_fnv1a.restype = C.c_uint64

    # -- culling ---------------------------------------------------------------------------
    def box_extent(self, lower4, upper4):
        ext = np.zeros_like(lower4)
        self.lib.dporacle_box_extent(_fp(lower4), _fp(upper4), len(lower4), _fp(ext))
        return ext

    def cull_bits(self, lower4, extent4, tidx, mats, vp, stride=64, threads=1):
        n = len(lower4)
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
        args = [_fp(lower4), _fp(extent4), _up(tidx), n, mats.ctypes.data, stride, _fp(vp), _up(words)]
        if threads > 1:
            self.lib.dporacle_cull_bits_mt(*args, threads)
        else:
            self.lib.dporacle_cull_bits(*args)
        return words

    def result_resize(self, words, old_n, new_n):
        need = (max(old_n, new_n) + 31) // 32
        if len(words) < need:
            words = np.concatenate([words, np.zeros(need - len(words), dtype=np.uint32)])
        self.lib.dporacle_result_resize(_up(words), old_n, new_n)
        return words[: (new_n + 31) // 32].copy()

    def update_changed(self, new_words, result_words, n):
        changed = np.empty(max(n, 1), dtype=np.uint32)
        cnt = self.lib.dporacle_update_changed(_up(new_words), _up(result_words), n, _up(changed))
        return changed[:cnt].copy()

    def result_move_bit(self, words, size, old, new):
        self.lib.dporacle_result_move_bit(_up(words), size, old, new)

    def bounding_box(self, lower4, extent4, tidx, mats, stride=64):
        out = np.zeros(6, dtype=np.float32)
        self.lib.dporacle_bounding_box(_fp(lower4), _fp(extent4), _up(tidx), len(lower4), mats.ctypes.data, stride, _fp(out))
        return out

    def fnv(self, words, n):
        return int(self.lib.dporacle_visibility_fnv1a(_up(words), n))

    # -- transform tree ---------------------------------------------------------------------
    def mat44_mul(self, a, b):
        r = np.zeros(16, dtype=np.float32)
        self.lib.dporacle_mat44_mul(_fp(np.ascontiguousarray(a, np.float32).reshape(16)),
                                    _fp(np.ascontiguousarray(b, np.float32).reshape(16)), _fp(r))
        return r.reshape(4, 4)

    def tree_compute(self, local, world, entries, level_offsets, dirty_local, dirty_world):
        """In-place on ``world``, ``dirty_local`` (cleared) and ``dirty_world`` (published set)."""
        n_nodes = local.shape[0]
        self.lib.dporacle_tree_compute(_fp(local.reshape(-1)), _fp(world.reshape(-1)), _up(entries.reshape(-1)),
                                       _up(level_offsets), len(level_offsets) - 1,
                                       _up(dirty_local), _up(dirty_world), n_nodes)


class RefCull:
    """One reference ``dp::culling::cpu::Manager`` + group (``kind = "reference"``)."""

    def __init__(self, lib, backend=0):
        self.lib = lib
        self.h = lib.dpref_cull_create(backend)
        if not self.h:
            raise RuntimeError("dpref_cull_create(%d) failed: %s" % (backend, lib.dpref_create_error().decode()))
        self._keep = None

    def close(self):
        if self.h:
            self.lib.dpref_cull_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.dpref_last_error(self.h).decode())

    def add_objects(self, lower3, upper3, tidx):
        lower3 = np.ascontiguousarray(lower3, np.float32)
        upper3 = np.ascontiguousarray(upper3, np.float32)
        tidx = np.ascontiguousarray(tidx, np.uint32)
        self._chk(self.lib.dpref_cull_add_objects(self.h, len(tidx), _fp(lower3.reshape(-1)), _fp(upper3.reshape(-1)), _up(tidx)))

    def set_object(self, index, lower3, upper3, tidx):
        l = np.ascontiguousarray(lower3, np.float32)
        u = np.ascontiguousarray(upper3, np.float32)
        self._chk(self.lib.dpref_cull_set_object(self.h, index, _fp(l), _fp(u), int(tidx)))

    def remove_object(self, index):
        self._chk(self.lib.dpref_cull_remove_object(self.h, index))

    def set_objects_many(self, indices, lower3, upper3, tidx):
        idx = np.ascontiguousarray(indices, np.uint32)
        l = np.ascontiguousarray(lower3, np.float32)
        u = np.ascontiguousarray(upper3, np.float32)
        t = np.ascontiguousarray(tidx, np.uint32)
        self._chk(self.lib.dpref_cull_set_objects_many(self.h, len(idx), _up(idx), _fp(l.reshape(-1)), _fp(u.reshape(-1)), _up(t)))

    def remove_objects_many(self, indices):
        idx = np.ascontiguousarray(indices, np.uint32)
        self._chk(self.lib.dpref_cull_remove_objects_many(self.h, len(idx), _up(idx)))

    def count(self):
        return int(self.lib.dpref_cull_count(self.h))

    def object_id(self, index):
        return int(self.lib.dpref_cull_object_id(self.h, index))

    def set_matrices(self, mats, stride=64, count=None):
        """``mats`` is BORROWED by the reference until the next cull; we keep a reference to it."""
        self._keep = mats
        if count is None:
            count = mats.nbytes // stride
        self._chk(self.lib.dpref_cull_set_matrices(self.h, mats.ctypes.data, count, stride))

    def matrix_changed(self, index):
        self._chk(self.lib.dpref_cull_matrix_changed(self.h, index))

    def matrices_changed(self, indices):
        idx = np.ascontiguousarray(indices, np.uint32)
        self._chk(self.lib.dpref_cull_matrices_changed(self.h, _up(idx), len(idx)))

    def result_create(self):
        r = self.lib.dpref_cull_result_create(self.h)
        if r < 0:
            self._chk(1)
        return r

    def cull(self, result, vp):
        vp = np.ascontiguousarray(vp, dtype=np.float32).reshape(16)
        self._chk(self.lib.dpref_cull_run(self.h, result, _fp(vp)))

    def changed(self, result):
        n = int(self.lib.dpref_cull_changed_count(self.h, result))
        out = np.empty(max(n, 1), dtype=np.uint32)
        self.lib.dpref_cull_changed(self.h, result, _up(out), n)
        return out[:n].copy()

    def changed_count(self, result):
        return int(self.lib.dpref_cull_changed_count(self.h, result))

    def visible_bits(self, result):
        n = self.count()
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        self._chk(self.lib.dpref_cull_visible_bits(self.h, result, _up(words)))
        return words

    def bounding_box(self):
        out = np.zeros(6, dtype=np.float32)
        self._chk(self.lib.dpref_cull_bounding_box(self.h, _fp(out)))
        return out

    # extensions of dp::culling::cuda::Manager (only in the drop-in test library)
    def set_device_matrices(self, device_ptr, count):
        self.lib.dpref_cull_set_device_matrices.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self._chk(self.lib.dpref_cull_set_device_matrices(self.h, device_ptr, count))

    def cull_multi(self, results, vps):
        vps = np.ascontiguousarray(vps, dtype=np.float32).reshape(-1)
        arr = (C.c_int * len(results))(*results)
        self.lib.dpref_cull_run_multi.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, _f32p]
        self._chk(self.lib.dpref_cull_run_multi(self.h, arr, len(results), _fp(vps)))


class RefTree:
    """One reference ``dp::transform::Tree``."""

    def __init__(self, lib, backend=0):
        """backend 0: dp::transform::Tree.  Drop-in test library only - 1: dp::transform::cuda::Tree edited
        through its own type, 2: the same class edited through a base-class reference."""
        self.lib = lib
        self.backend = backend
        if backend:
            self.h = lib.dpref_tree_create_backend(backend)
            if not self.h:
                raise RuntimeError(lib.dpref_create_error().decode(errors="replace"))
        else:
            self.h = lib.dpref_tree_create()

    def close(self):
        if self.h:
            self.lib.dpref_tree_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, parent, local):
        m = np.ascontiguousarray(local, np.float32).reshape(16)
        r = int(self.lib.dpref_tree_add(self.h, int(parent), _fp(m)))
        if r < 0:
            raise RuntimeError("Tree::addTransform failed")
        return r

    def add_many(self, parents, locals16):
        parents = np.ascontiguousarray(parents, np.uint32)
        locals16 = np.ascontiguousarray(locals16, np.float32).reshape(-1)
        out = np.empty(len(parents), dtype=np.uint32)
        if self.lib.dpref_tree_add_many(self.h, len(parents), _up(parents), _fp(locals16), _up(out)) != 0:
            raise RuntimeError("Tree::addTransform failed")
        return out

    def remove(self, index):
        if self.lib.dpref_tree_remove(self.h, int(index)) != 0:
            raise RuntimeError("Tree::removeTransform failed")

    def update_locals(self, indices, locals16):
        indices = np.ascontiguousarray(indices, np.uint32)
        locals16 = np.ascontiguousarray(locals16, np.float32).reshape(-1)
        self.lib.dpref_tree_update_locals(self.h, len(indices), _up(indices), _fp(locals16))

    def compute(self):
        if self.lib.dpref_tree_compute(self.h) != 0:
            raise RuntimeError(self.lib.dpref_create_error().decode(errors="replace"))

    def device_world(self):
        """device pointer of the world matrices (dp::transform::cuda::Tree only)"""
        return int(self.lib.dpref_tree_device_world(self.h) or 0)

    def host_mirror(self, enable):
        self.lib.dpref_tree_host_mirror(self.h, int(bool(enable)))

    def count(self):
        return int(self.lib.dpref_tree_count(self.h))

    def world(self):
        """Copy of the world-matrix array, shape (count, 4, 4)."""
        n = self.count()
        p = self.lib.dpref_tree_world(self.h)
        return np.ctypeslib.as_array(p, shape=(n * 16,)).reshape(n, 4, 4).copy()

    def world_view(self):
        """Borrowed view of the tree's own storage (what CullingImpl hands to groupSetMatrices)."""
        n = self.count()
        p = self.lib.dpref_tree_world(self.h)
        return np.ctypeslib.as_array(p, shape=(n * 16,))

    def dirty_world(self):
        n = self.count()
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        self.lib.dpref_tree_dirty_world(self.h, _up(words), len(words))
        return words


def _declare_ref(lib):
    lib.dpref_cull_create.argtypes = [C.c_int]
    lib.dpref_cull_create.restype = C.c_void_p
    lib.dpref_cull_destroy.argtypes = [C.c_void_p]
    lib.dpref_create_error.restype = C.c_char_p
    lib.dpref_last_error.argtypes = [C.c_void_p]
    lib.dpref_last_error.restype = C.c_char_p
    lib.dpref_cull_add_objects.argtypes = [C.c_void_p, C.c_size_t, _f32p, _f32p, _u32p]
    lib.dpref_cull_set_object.argtypes = [C.c_void_p, C.c_size_t, _f32p, _f32p, C.c_uint32]
    lib.dpref_cull_remove_object.argtypes = [C.c_void_p, C.c_size_t]
    if hasattr(lib, "dpref_cull_set_objects_many"):
        lib.dpref_cull_set_objects_many.argtypes = [C.c_void_p, C.c_size_t, _u32p, _f32p, _f32p, _u32p]
        lib.dpref_cull_remove_objects_many.argtypes = [C.c_void_p, C.c_size_t, _u32p]
    lib.dpref_cull_count.argtypes = [C.c_void_p]
    lib.dpref_cull_count.restype = C.c_size_t
    lib.dpref_cull_object_id.argtypes = [C.c_void_p, C.c_size_t]
    lib.dpref_cull_object_id.restype = C.c_uint32
    lib.dpref_cull_set_matrices.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.dpref_cull_matrix_changed.argtypes = [C.c_void_p, C.c_size_t]
    lib.dpref_cull_matrices_changed.argtypes = [C.c_void_p, _u32p, C.c_size_t]
    lib.dpref_cull_result_create.argtypes = [C.c_void_p]
    lib.dpref_cull_run.argtypes = [C.c_void_p, C.c_int, _f32p]
    lib.dpref_cull_changed_count.argtypes = [C.c_void_p, C.c_int]
    lib.dpref_cull_changed_count.restype = C.c_size_t
    lib.dpref_cull_changed.argtypes = [C.c_void_p, C.c_int, _u32p, C.c_size_t]
    lib.dpref_cull_changed.restype = C.c_size_t
    lib.dpref_cull_visible_bits.argtypes = [C.c_void_p, C.c_int, _u32p]
    lib.dpref_cull_bounding_box.argtypes = [C.c_void_p, _f32p]
    lib.dpref_tree_create.restype = C.c_void_p
    lib.dpref_tree_destroy.argtypes = [C.c_void_p]
    lib.dpref_tree_add.argtypes = [C.c_void_p, C.c_uint32, _f32p]
    lib.dpref_tree_add.restype = C.c_int64
    lib.dpref_tree_add_many.argtypes = [C.c_void_p, C.c_size_t, _u32p, _f32p, _u32p]
    lib.dpref_tree_remove.argtypes = [C.c_void_p, C.c_uint32]
    lib.dpref_tree_update_local.argtypes = [C.c_void_p, C.c_uint32, _f32p]
    lib.dpref_tree_update_locals.argtypes = [C.c_void_p, C.c_size_t, _u32p, _f32p]
    lib.dpref_tree_compute.argtypes = [C.c_void_p]
    lib.dpref_tree_compute.restype = C.c_int
    if hasattr(lib, "dpref_tree_create_backend"):          # the drop-in test library (tests/cpp)
        lib.dpref_tree_create_backend.argtypes = [C.c_int]
        lib.dpref_tree_create_backend.restype = C.c_void_p
        lib.dpref_tree_device_world.argtypes = [C.c_void_p]
        lib.dpref_tree_device_world.restype = C.c_void_p
        lib.dpref_tree_host_mirror.argtypes = [C.c_void_p, C.c_int]
    lib.dpref_tree_count.argtypes = [C.c_void_p]
    lib.dpref_tree_count.restype = C.c_size_t
    lib.dpref_tree_world.argtypes = [C.c_void_p]
    lib.dpref_tree_world.restype = _f32p
    lib.dpref_tree_dirty_world.argtypes = [C.c_void_p, _u32p, C.c_size_t]
    lib.dpref_tree_dirty_world.restype = C.c_size_t
    return lib


class Reference:
    """The compiled, unmodified reference.  ``Reference.available()`` is False when the
    prebuilt library did not travel (then tests that need it skip and say so)."""

    def __init__(self, path: str | None = None):
        path = path or REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = _declare_ref(C.CDLL(path))

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def cull(self, backend=0) -> RefCull:
        return RefCull(self.lib, backend)

    def tree(self, backend=0) -> RefTree:
        return RefTree(self.lib, backend)
