"""ctypes binding of include/dpcu.h (pipeline_b200/lib/libdpcu.so).

This is plumbing for the Python test-suite and bench.py; the product is the shared library
and the C++ ``dp::culling::cuda::Manager`` on top of it.  Loading fails loudly when the
library has not been built - there is no fallback of any kind.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# DPCU_LIB: developer override (tools/ builds tuning variants of the library next to the product one)
LIB_PATH = os.environ.get("DPCU_LIB") or os.path.join(HERE, "lib", "libdpcu.so")

MEM_HOST, MEM_DEVICE = 0, 1
OPT_KERNEL, OPT_FMA, OPT_CHANGED_LIST, OPT_CTAS_PER_SM, OPT_PROFILE, OPT_FUSE_LEAF, OPT_FUSE_LIST, OPT_LAST_KERNEL, OPT_FILTER, OPT_LINE_WORDS, OPT_LIST_OFFSETS, OPT_L2_PREFETCH = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12
TREE_OPT_WIDE_MIN_NODES = 1
KERNEL_NAMES = {1: "cullDirectKernel", 2: "cullStagedKernel", 3: "cullViewsKernel", 4: "cullLinesKernel", 5: "cullViewsKernel", 6: "cullFusedLeafKernel",
                7: "cullLinesMvKernel", 8: "cullGridKernel"}
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_STAGED, KERNEL_VIEWS, KERNEL_LINES, KERNEL_VIEWS_CHAINS, KERNEL_FUSED_LEAF, KERNEL_LINES_PAIRS, KERNEL_GRID = 0, 1, 2, 3, 4, 5, 6, 7, 8
MAX_VIEWS = 8

_vp = C.c_void_p
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_szp = C.POINTER(C.c_size_t)

_lib = None

# Every wrapper that owns a library handle registers here; at interpreter exit the handles are
# released in dependency order (results before their context, contexts before buffers) while
# the CUDA runtime is still alive, and finalisers that run later become no-ops.
_live = weakref.WeakSet()
_shutdown = False
_CLOSE_ORDER = ("CullResult", "Cull", "Tree", "Event", "Stream", "HostBuffer", "Buffer")


def _register(obj):
    _live.add(obj)


@atexit.register
def _close_all():
    global _shutdown
    objs = list(_live)
    for name in _CLOSE_ORDER:
        for o in objs:
            if type(o).__name__ == name:
                try:
                    o.close()
                except Exception:
                    pass
    _shutdown = True


class DpcuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dpcu error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C pipeline_b200/csrc); there is no fallback path" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.dpcuGetLastError.restype = C.c_char_p
        _declare(L)
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise DpcuError(rc, lib().dpcuGetLastError().decode(errors="replace"))


def _declare(L):
    sig = {
        "dpcuDeviceCount": [C.POINTER(C.c_int)],
        "dpcuDeviceSelect": [C.c_int],
        "dpcuDeviceCurrent": [C.POINTER(C.c_int)],
        "dpcuDeviceSynchronize": [],
        "dpcuDeviceInfo": [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), _szp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "dpcuDeviceEnablePeerAccess": [C.c_int, C.c_int],
        "dpcuBufferCreate": [C.POINTER(_vp), C.c_size_t],
        "dpcuBufferDestroy": [_vp],
        "dpcuBufferSize": [_vp, _szp],
        "dpcuBufferDevicePointer": [_vp, C.POINTER(_vp)],
        "dpcuBufferUpload": [_vp, C.c_size_t, _vp, C.c_size_t, _vp],
        "dpcuBufferDownload": [_vp, C.c_size_t, _vp, C.c_size_t, _vp],
        "dpcuBufferFill": [_vp, C.c_int, C.c_size_t, C.c_size_t],
        "dpcuBufferFillAsync": [_vp, C.c_int, C.c_size_t, C.c_size_t, _vp],
        "dpcuHostBufferCreate": [C.POINTER(_vp), C.c_size_t, C.c_uint],
        "dpcuHostBufferDestroy": [_vp],
        "dpcuHostBufferPointer": [_vp, C.POINTER(_vp)],
        "dpcuHostBufferSize": [_vp, _szp],
        "dpcuStreamCreate": [C.POINTER(_vp), C.c_int, C.c_int],
        "dpcuStreamDestroy": [_vp],
        "dpcuStreamSynchronize": [_vp],
        "dpcuStreamIsCompleted": [_vp, C.POINTER(C.c_int)],
        "dpcuStreamWaitEvent": [_vp, _vp],
        "dpcuStreamNative": [_vp, C.POINTER(_vp)],
        "dpcuEventCreate": [C.POINTER(_vp), C.c_uint],
        "dpcuEventDestroy": [_vp],
        "dpcuEventRecord": [_vp, _vp],
        "dpcuEventSynchronize": [_vp],
        "dpcuEventIsCompleted": [_vp, C.POINTER(C.c_int)],
        "dpcuEventElapsedMs": [_vp, _vp, C.POINTER(C.c_float)],
        "dpcuCullCreate": [C.POINTER(_vp), C.c_int],
        "dpcuCullDestroy": [_vp],
        "dpcuCullSetObjects": [_vp, _vp, _vp, _vp, C.c_size_t, C.c_int],
        "dpcuCullSetObjectRange": [_vp, C.c_size_t, C.c_size_t, _vp, _vp, _vp, C.c_int],
        "dpcuCullGetObjectCount": [_vp, _szp],
        "dpcuCullSetMatrices": [_vp, _vp, C.c_size_t, C.c_size_t, C.c_int],
        "dpcuCullUpdateMatrices": [_vp, _u32p, C.c_size_t, _vp, C.c_size_t, C.c_int],
        "dpcuCullBindMatrices": [_vp, _vp, C.c_size_t],
        "dpcuCullBindTree": [_vp, _vp],
        "dpcuCullSetObjectCount": [_vp, C.c_size_t],
        "dpcuCullUpdateObjects": [_vp, _u32p, C.c_size_t, _vp, _vp, _u32p],
        "dpcuCullResultUpdateWords": [_vp, _u32p, _u32p, C.c_size_t],
        "dpcuTreeGetWorldDirty": [_vp, _vp, C.c_size_t, _szp],
        "dpcuCullGetMatrixCount": [_vp, _szp],
        "dpcuCullResultCreate": [_vp, C.POINTER(_vp)],
        "dpcuCullResultDestroy": [_vp],
        "dpcuCullRun": [_vp, C.POINTER(_vp), _f32p, C.c_int, _vp],
        "dpcuCullRunWithTree": [_vp, _vp, C.POINTER(_vp), _f32p, C.c_int, _vp],
        "dpcuCullResultGetBits": [_vp, _u32p, C.c_size_t],
        "dpcuCullResultGetChangedCount": [_vp, _szp],
        "dpcuCullResultGetChanged": [_vp, _u32p, C.c_size_t, _szp],
        "dpcuCullResultIsVisible": [_vp, C.c_size_t, C.POINTER(C.c_int)],
        "dpcuCullResultMoveBit": [_vp, C.c_size_t, C.c_size_t],
        "dpcuCullResultDevicePointers": [_vp, C.POINTER(_vp), _szp, C.POINTER(_vp), C.POINTER(_vp)],
        "dpcuCullGetBoundingBox": [_vp, _f32p],
        "dpcuCullSetOption": [_vp, C.c_int, C.c_int],
        "dpcuCullGetOption": [_vp, C.c_int, C.POINTER(C.c_int)],
        "dpcuCullGetLaunchCount": [_vp, C.POINTER(C.c_uint64)],
        "dpcuDebugKernelArgLayout": [C.c_int, _szp, _szp, _szp, _szp],
        "dpcuCullGetKernelTime": [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)],
        "dpcuCullGetKernelTimes": [_vp, _f32p, C.c_size_t, _szp],
        "dpcuCullResultSetPeerBits": [_vp, C.POINTER(_vp), C.c_int, C.c_size_t],
        "dpcuCullResultBuildVisibleList": [_vp, _vp],
        "dpcuCullResultVisibleDevicePointers": [_vp, C.POINTER(_vp), C.POINTER(_vp)],
        "dpcuCullResultGetVisible": [_vp, _u32p, C.c_size_t, _szp],
        "dpcuCullResultSetHostMirror": [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp],
        "dpcuCullResultSynchronize": [_vp],
        "dpcuIpcGetHandle": [_vp, C.c_char_p],
        "dpcuIpcOpen": [C.c_char_p, C.POINTER(_vp)],
        "dpcuIpcClose": [_vp],
        "dpcuTreeCreate": [C.POINTER(_vp), C.c_int],
        "dpcuTreeDestroy": [_vp],
        "dpcuTreeSetTopology": [_vp, _u32p, _u32p, C.c_int, C.c_size_t],
        "dpcuTreeSetLocals": [_vp, C.c_size_t, C.c_size_t, _vp, C.c_int],
        "dpcuTreeUpdateLocals": [_vp, _u32p, C.c_size_t, _vp, C.c_int],
        "dpcuTreeMarkDirty": [_vp, C.c_size_t, C.c_size_t],
        "dpcuTreeCompute": [_vp, _vp],
        "dpcuTreeWorldDevicePointer": [_vp, C.POINTER(_vp), _szp],
        "dpcuTreeLocalDevicePointer": [_vp, C.POINTER(_vp), _szp],
        "dpcuTreeGetWorld": [_vp, C.c_size_t, C.c_size_t, _vp],
        "dpcuTreeGetDirtyWorld": [_vp, _u32p, C.c_size_t],
        "dpcuTreeGetLaunchCount": [_vp, C.POINTER(C.c_uint64)],
        "dpcuTreeSetOption": [_vp, C.c_int, C.c_size_t],
        "dpcuSceneGenerate": [C.c_uint64, C.c_uint64, C.c_size_t, C.c_uint32, _vp, _vp, _vp, _vp],
        "dpcuDebugReadSweep": [_vp, C.c_size_t, _vp],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    L.dpcuGetVersion.restype = C.c_int


EXPORTED = None  # filled by tests from include/dpcu.h


def _ptr(a):
    """numpy array -> void*, int -> void* (device pointer), None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)


# ------------------------------------------------------------------------------ dp/cuda layer
def device_count():
    n = C.c_int(0)
    check(lib().dpcuDeviceCount(C.byref(n)))
    return n.value


def device_select(i):
    check(lib().dpcuDeviceSelect(i))


def device_sync():
    check(lib().dpcuDeviceSynchronize())


def device_info(i=0):
    name = C.create_string_buffer(256)
    sm, mem, maj, mnr = C.c_int(), C.c_size_t(), C.c_int(), C.c_int()
    check(lib().dpcuDeviceInfo(i, name, 256, C.byref(sm), C.byref(mem), C.byref(maj), C.byref(mnr)))
    return {"name": name.value.decode(), "sms": sm.value, "mem": mem.value, "cc": (maj.value, mnr.value)}


class Buffer:
    """dp::cuda::Buffer"""

    def __init__(self, nbytes):
        self.h = _vp()
        check(lib().dpcuBufferCreate(C.byref(self.h), nbytes))
        _register(self)
        self.nbytes = nbytes

    @property
    def ptr(self):
        p = _vp()
        check(lib().dpcuBufferDevicePointer(self.h, C.byref(p)))
        return p.value or 0

    def upload(self, arr, offset=0, stream=None):
        check(lib().dpcuBufferUpload(self.h, offset, _ptr(arr), arr.nbytes, stream.h if stream else None))

    def download(self, arr, offset=0, stream=None):
        check(lib().dpcuBufferDownload(self.h, offset, _ptr(arr), arr.nbytes, stream.h if stream else None))
        return arr

    def fill(self, byte, stream=None, nbytes=None, offset=0):
        n = self.nbytes if nbytes is None else nbytes
        if stream is not None:
            check(lib().dpcuBufferFillAsync(self.h, byte, n, offset, stream.h))
        else:
            check(lib().dpcuBufferFill(self.h, byte, n, offset))

    def close(self):
        if self.h:
            check(lib().dpcuBufferDestroy(self.h))
            self.h = None

    def __del__(self):
        if _shutdown:
            return
        try:
            self.close()
        except Exception:
            pass


class HostBuffer:
    """dp::cuda::BufferHost (pinned)."""

    def __init__(self, nbytes, flags=0):
        self.h = _vp()
        check(lib().dpcuHostBufferCreate(C.byref(self.h), nbytes, flags))
        _register(self)
        self.nbytes = nbytes
        p = _vp()
        check(lib().dpcuHostBufferPointer(self.h, C.byref(p)))
        self.ptr = p.value or 0

    def array(self, dtype, count=None, offset=0):
        dtype = np.dtype(dtype)
        if count is None:
            count = (self.nbytes - offset) // dtype.itemsize
        buf = (C.c_char * (count * dtype.itemsize)).from_address(self.ptr + offset)
        return np.frombuffer(buf, dtype=dtype, count=count)

    def close(self):
        if self.h:
            check(lib().dpcuHostBufferDestroy(self.h))
            self.h = None

    def __del__(self):
        if _shutdown:
            return
        try:
            self.close()
        except Exception:
            pass


class Stream:
    def __init__(self, blocking=False, priority=0):
        self.h = _vp()
        check(lib().dpcuStreamCreate(C.byref(self.h), int(blocking), priority))
        _register(self)

    def sync(self):
        check(lib().dpcuStreamSynchronize(self.h))

    def completed(self):
        c = C.c_int()
        check(lib().dpcuStreamIsCompleted(self.h, C.byref(c)))
        return bool(c.value)

    def wait(self, event):
        check(lib().dpcuStreamWaitEvent(self.h, event.h))

    @property
    def native(self):
        p = _vp()
        check(lib().dpcuStreamNative(self.h, C.byref(p)))
        return p.value or 0

    def close(self):
        if self.h:
            check(lib().dpcuStreamDestroy(self.h))
            self.h = None

    def __del__(self):
        if _shutdown:
            return
        try:
            self.close()
        except Exception:
            pass


class Event:
    def __init__(self, flags=0):
        self.h = _vp()
        check(lib().dpcuEventCreate(C.byref(self.h), flags))
        _register(self)

    def record(self, stream=None):
        check(lib().dpcuEventRecord(self.h, stream.h if stream else None))

    def sync(self):
        check(lib().dpcuEventSynchronize(self.h))

    def elapsed_ms(self, stop):
        ms = C.c_float()
        check(lib().dpcuEventElapsedMs(self.h, stop.h, C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            check(lib().dpcuEventDestroy(self.h))
            self.h = None

    def __del__(self):
        if _shutdown:
            return
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------ culling layer
class CullResult:
    def __init__(self, ctx):
        self.ctx = ctx
        self.h = _vp()
        check(lib().dpcuCullResultCreate(ctx.h, C.byref(self.h)))
        _register(self)

    def bits(self):
        n = self.ctx.count()
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        check(lib().dpcuCullResultGetBits(self.h, words.ctypes.data_as(_u32p), len(words)))
        return words

    def changed_count(self):
        c = C.c_size_t()
        check(lib().dpcuCullResultGetChangedCount(self.h, C.byref(c)))
        return c.value

    def changed(self):
        n = self.changed_count()
        out = np.empty(max(n, 1), dtype=np.uint32)
        c = C.c_size_t()
        check(lib().dpcuCullResultGetChanged(self.h, out.ctypes.data_as(_u32p), n, C.byref(c)))
        return out[:n].copy()

    def is_visible(self, index):
        v = C.c_int()
        check(lib().dpcuCullResultIsVisible(self.h, index, C.byref(v)))
        return bool(v.value)

    def move_bit(self, old, new):
        check(lib().dpcuCullResultMoveBit(self.h, old, new))

    def update_words(self, indices, words):
        """overwrite visibility words of the stored result (a frame's bit moves applied on a host mirror, in one batch)"""
        idx = np.ascontiguousarray(indices, np.uint32)
        w = np.ascontiguousarray(words, np.uint32)
        check(lib().dpcuCullResultUpdateWords(self.h, idx.ctypes.data_as(_u32p), w.ctypes.data_as(_u32p), len(idx)))

    def device_pointers(self):
        bits, chg, cnt, nw = _vp(), _vp(), _vp(), C.c_size_t()
        check(lib().dpcuCullResultDevicePointers(self.h, C.byref(bits), C.byref(nw), C.byref(chg), C.byref(cnt)))
        return {"bits": bits.value or 0, "n_words": nw.value, "changed": chg.value or 0, "count": cnt.value or 0}

    def set_peer_bits(self, pointers, word_offset):
        arr = (_vp * max(len(pointers), 1))(*[(_vp(p) if p else None) for p in pointers])
        check(lib().dpcuCullResultSetPeerBits(self.h, arr, len(pointers), word_offset))

    def build_visible_list(self, stream=None):
        check(lib().dpcuCullResultBuildVisibleList(self.h, stream.h if stream else None))

    def visible_device_pointers(self):
        idx, cnt = _vp(), _vp()
        check(lib().dpcuCullResultVisibleDevicePointers(self.h, C.byref(idx), C.byref(cnt)))
        return idx.value or 0, cnt.value or 0

    def visible(self):
        c = C.c_size_t()
        check(lib().dpcuCullResultGetVisible(self.h, None, 0, C.byref(c)))
        out = np.empty(max(c.value, 1), dtype=np.uint32)
        check(lib().dpcuCullResultGetVisible(self.h, out.ctypes.data_as(_u32p), c.value, C.byref(c)))
        return out[:c.value].copy()

    def set_host_mirror(self, bits=None, changed=None, count=None):
        """bits / changed / count: uint32 numpy views of pinned HostBuffer memory (or None).  After
        run() + synchronize() they hold the visibility words, the changed list and its length."""
        self._mirror = (bits, changed, count)          # keep the views (and their buffers) alive
        check(lib().dpcuCullResultSetHostMirror(
            self.h, _ptr(bits), 0 if bits is None else bits.size, _ptr(changed), 0 if changed is None else changed.size, _ptr(count)))

    def synchronize(self):
        check(lib().dpcuCullResultSynchronize(self.h))

    def close(self):
        if self.h:
            check(lib().dpcuCullResultDestroy(self.h))
            self.h = None
            self._mirror = None

    def __del__(self):
        if _shutdown:
            return
        try:
            self.close()
        except Exception:
            pass


class Cull:
    """One culling group on one device (dpcuCull)."""

    def __init__(self, device=0):
        self.h = _vp()
        check(lib().dpcuCullCreate(C.byref(self.h), device))
        _register(self)
        self.device = device
        self._mat_src = None

    def set_objects(self, lower4, extent4, tidx, memspace=MEM_HOST, n=None):
        if n is None:
            n = len(lower4)
        check(lib().dpcuCullSetObjects(self.h, _ptr(lower4), _ptr(extent4), _ptr(tidx), n, memspace))

    def set_object_range(self, first, lower4, extent4, tidx, memspace=MEM_HOST, count=None):
        if count is None:
            count = len(lower4)
        check(lib().dpcuCullSetObjectRange(self.h, first, count, _ptr(lower4), _ptr(extent4), _ptr(tidx), memspace))

    def count(self):
        c = C.c_size_t()
        check(lib().dpcuCullGetObjectCount(self.h, C.byref(c)))
        return c.value

    def set_object_count(self, n):
        """grow / shrink the object arrays keeping their contents"""
        check(lib().dpcuCullSetObjectCount(self.h, n))

    def update_objects(self, indices, lower4, extent4, tidx):
        """one batch of object edits: object indices[k] <- (lower4[k], extent4[k], tidx[k])"""
        idx = np.ascontiguousarray(indices, np.uint32)
        lo = np.ascontiguousarray(lower4, np.float32)
        ex = np.ascontiguousarray(extent4, np.float32)
        ti = np.ascontiguousarray(tidx, np.uint32)
        check(lib().dpcuCullUpdateObjects(self.h, idx.ctypes.data_as(_u32p), len(idx), _ptr(lo), _ptr(ex), ti.ctypes.data_as(_u32p)))

    def set_matrices(self, mats, stride=64, count=None, memspace=MEM_HOST):
        if count is None:
            count = mats.nbytes // stride
        check(lib().dpcuCullSetMatrices(self.h, _ptr(mats), count, stride, memspace))
        if memspace == MEM_HOST:
            self._mat_src = (mats, stride)

    def update_matrices(self, indices, mats, stride=64):
        idx = np.ascontiguousarray(indices, np.uint32)
        check(lib().dpcuCullUpdateMatrices(self.h, idx.ctypes.data_as(_u32p), len(idx), _ptr(mats), stride, MEM_HOST))

    def update_matrices_from_source(self, indices):
        mats, stride = self._mat_src
        self.update_matrices(indices, mats, stride)

    def bind_matrices(self, device_ptr, count):
        check(lib().dpcuCullBindMatrices(self.h, device_ptr, count))

    def fma_report(self, vp, limit=16):
        """north_star: "any fused-multiply-add fast mode must report its boundary-object disagreements separately".
        Culls the resident scene against `vp` twice into fresh results - exact (-fmad=false) and the reporting-only
        FMA build of the direct kernel - and returns the objects whose visibility differs."""
        vp = np.ascontiguousarray(vp, np.float32).reshape(1, 16)
        exact, fast = self.result_create(), self.result_create()
        keep = [self.get_option(o) for o in (OPT_FMA, OPT_KERNEL)]
        try:
            self.set_option(OPT_FMA, 0)
            self.run([exact], vp)
            self.set_option(OPT_FMA, 1)
            self.set_option(OPT_KERNEL, KERNEL_AUTO)
            self.run([fast], vp)
            diff = exact.bits() ^ fast.bits()
        finally:
            self.set_option(OPT_FMA, keep[0])
            self.set_option(OPT_KERNEL, keep[1])
            exact.close(), fast.close()
        words = np.flatnonzero(diff)
        idx = []
        for w in words:
            x = int(diff[w])
            while x:
                b = (x & -x).bit_length() - 1
                idx.append(int(w) * 32 + b)
                x &= x - 1
        return {"objects": int(self.count()), "disagreements": len(idx), "first_indices": idx[:limit], "indices": idx}

    def bind_tree(self, tree):
        """cull out of the tree's world matrices in place; culls and tree computes are ordered by events"""
        check(lib().dpcuCullBindTree(self.h, tree.h))

    def matrix_count(self):
        c = C.c_size_t()
        check(lib().dpcuCullGetMatrixCount(self.h, C.byref(c)))
        return c.value

    def result_create(self):
        return CullResult(self)

    def run(self, results, vps, stream=None):
        vps = np.ascontiguousarray(vps, dtype=np.float32).reshape(-1)
        nv = len(results)
        assert len(vps) == 16 * nv
        arr = (_vp * nv)(*[r.h for r in results])
        check(lib().dpcuCullRun(self.h, arr, vps.ctypes.data_as(_f32p), nv, stream.h if stream else None))

    def run_with_tree(self, tree, results, vps, stream=None):
        """Tree::compute + cull in one call (the tree's last level fused into the cull kernel when
        object i is bound to that level's entry i)."""
        vps = np.ascontiguousarray(vps, dtype=np.float32).reshape(-1)
        nv = len(results)
        assert len(vps) == 16 * nv
        arr = (_vp * nv)(*[r.h for r in results])
        check(lib().dpcuCullRunWithTree(self.h, tree.h, arr, vps.ctypes.data_as(_f32p), nv, stream.h if stream else None))

    def bounding_box(self):
        out = np.zeros(6, dtype=np.float32)
        check(lib().dpcuCullGetBoundingBox(self.h, out.ctypes.data_as(_f32p)))
        return out

    def set_option(self, opt, value):
        check(lib().dpcuCullSetOption(self.h, opt, value))

    def get_option(self, opt):
        v = C.c_int()
        check(lib().dpcuCullGetOption(self.h, opt, C.byref(v)))
        return v.value

    def launches(self):
        v = C.c_uint64()
        check(lib().dpcuCullGetLaunchCount(self.h, C.byref(v)))
        return v.value

    def kernel_time(self):
        """(total ms, launches) of the cull kernel since the last call; needs OPT_PROFILE = 1."""
        ms, n = C.c_double(), C.c_uint64()
        check(lib().dpcuCullGetKernelTime(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def kernel_times(self, capacity=4096):
        """per-launch ms of the cull kernel since the last call (oldest first); needs OPT_PROFILE = 1."""
        out = np.zeros(capacity, np.float32)
        n = C.c_size_t()
        check(lib().dpcuCullGetKernelTimes(self.h, out.ctypes.data_as(_f32p), capacity, C.byref(n)))
        return out[:min(n.value, capacity)].copy()

    def close(self):
        if self.h:
            check(lib().dpcuCullDestroy(self.h))
            self.h = None

    def __del__(self):
        if _shutdown:
            return
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------ transform layer
class Tree:
    def __init__(self, device=0):
        self.h = _vp()
        check(lib().dpcuTreeCreate(C.byref(self.h), device))
        _register(self)
        self.n_nodes = 0

    def set_topology(self, entries, level_offsets, n_nodes):
        entries = np.ascontiguousarray(entries, np.uint32).reshape(-1)
        level_offsets = np.ascontiguousarray(level_offsets, np.uint32)
        check(lib().dpcuTreeSetTopology(self.h, entries.ctypes.data_as(_u32p), level_offsets.ctypes.data_as(_u32p),
                                        len(level_offsets) - 1, n_nodes))
        self.n_nodes = n_nodes

    def set_locals(self, first, mats, memspace=MEM_HOST, count=None):
        if count is None:
            count = mats.nbytes // 64
        check(lib().dpcuTreeSetLocals(self.h, first, count, _ptr(mats), memspace))

    def update_locals(self, indices, mats):
        idx = np.ascontiguousarray(indices, np.uint32)
        mats = np.ascontiguousarray(mats, np.float32)
        check(lib().dpcuTreeUpdateLocals(self.h, idx.ctypes.data_as(_u32p), len(idx), _ptr(mats), MEM_HOST))

    def mark_dirty(self, first, count):
        check(lib().dpcuTreeMarkDirty(self.h, first, count))

    def compute(self, stream=None):
        check(lib().dpcuTreeCompute(self.h, stream.h if stream else None))

    def world_ptr(self):
        p, n = _vp(), C.c_size_t()
        check(lib().dpcuTreeWorldDevicePointer(self.h, C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def local_ptr(self):
        p, n = _vp(), C.c_size_t()
        check(lib().dpcuTreeLocalDevicePointer(self.h, C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def world(self, first=0, count=None):
        if count is None:
            count = self.n_nodes - first
        out = np.zeros((count, 4, 4), dtype=np.float32)
        check(lib().dpcuTreeGetWorld(self.h, first, count, _ptr(out)))
        return out

    def refresh_host_world(self, host_world):
        """scatter the world matrices the last compute changed into host_world ((n_nodes, 4, 4) float32); returns how many"""
        assert host_world.flags["C_CONTIGUOUS"] and host_world.dtype == np.float32
        c = C.c_size_t()
        check(lib().dpcuTreeGetWorldDirty(self.h, host_world.ctypes.data, host_world.size // 16, C.byref(c)))
        return c.value

    def dirty_world(self):
        words = np.zeros((self.n_nodes + 31) // 32, dtype=np.uint32)
        check(lib().dpcuTreeGetDirtyWorld(self.h, words.ctypes.data_as(_u32p), len(words)))
        return words

    def set_option(self, option, value):
        check(lib().dpcuTreeSetOption(self.h, option, value))

    def launches(self):
        v = C.c_uint64()
        check(lib().dpcuTreeGetLaunchCount(self.h, C.byref(v)))
        return v.value

    def close(self):
        if self.h:
            check(lib().dpcuTreeDestroy(self.h))
            self.h = None

    def __del__(self):
        if _shutdown:
            return
        try:
            self.close()
        except Exception:
            pass


def ipc_get_handle(device_ptr) -> bytes:
    buf = C.create_string_buffer(64)
    check(lib().dpcuIpcGetHandle(device_ptr, buf))
    return buf.raw


def ipc_open(handle: bytes) -> int:
    p = _vp()
    check(lib().dpcuIpcOpen(handle, C.byref(p)))
    return p.value or 0


def ipc_close(device_ptr):
    check(lib().dpcuIpcClose(device_ptr))


def enable_peer_access(device, peer):
    check(lib().dpcuDeviceEnablePeerAccess(device, peer))


def scene_generate(seed, first, count, index_base, lower_ptr, extent_ptr, mats_ptr, stream=None):
    check(lib().dpcuSceneGenerate(seed, first, count, index_base, lower_ptr, extent_ptr, mats_ptr,
                                  stream.h if stream else None))


def read_sweep(device_ptr, nbytes, stream=None):
    """bench support: read a device buffer once (leaves the L2 cold and clean after a flush write)"""
    check(lib().dpcuDebugReadSweep(device_ptr, nbytes, stream.h if stream else None))
