"""Three interchangeable drivers of the culling path, used by the parity tests:

* ``RefEngine``  - the compiled unmodified reference (oracle/_ref/libdpref.so),
* ``PortEngine`` - the C restatement (oracle/libdporacle.so) plus the group / result
  bookkeeping the reference keeps on the host (GroupBitSet / ResultBitSet),
* ``CudaEngine`` - the product: pipeline_b200's C-ABI (needs a GPU).

All expose: add(lower4, upper4, tidx) / remove(group_index) / set_matrices(raw, stride) /
matrices_changed(indices) / cull(vp) -> (bits words, changed indices) / bounding_box() / count().
"""
from __future__ import annotations

import numpy as np

from oracle.loader import Port, Reference


class RefEngine:
    kind = "reference"

    def __init__(self, ref: Reference | None = None, backend: int = 0):
        self.ref = ref or Reference()
        self.s = self.ref.cull(backend)
        self.r = self.s.result_create()
        if backend:
            self.kind = "dp::culling::cuda::Manager"

    def set_object(self, index, lower4, upper4, tidx):
        self.s.set_object(index, lower4[:3], upper4[:3], tidx)

    def add(self, lower4, upper4, tidx):
        self.s.add_objects(np.ascontiguousarray(lower4[:, :3]), np.ascontiguousarray(upper4[:, :3]), tidx)

    def remove(self, index):
        self.s.remove_object(index)

    def set_matrices(self, raw, stride=64, count=None):
        self.s.set_matrices(raw, stride, count)

    def matrices_changed(self, indices):
        self.s.matrices_changed(indices)

    def cull(self, vp):
        self.s.cull(self.r, vp)
        return self.s.visible_bits(self.r), self.s.changed(self.r)

    def bounding_box(self):
        return self.s.bounding_box()

    def count(self):
        return self.s.count()

    def close(self):
        self.s.close()


class PortEngine:
    kind = "port"

    def __init__(self, port: Port | None = None):
        self.p = port or Port()
        self.lower4 = np.zeros((0, 4), np.float32)
        self.extent4 = np.zeros((0, 4), np.float32)
        self.tidx = np.zeros(0, np.uint32)
        self.mats = None
        self.stride = 64
        self.result = np.zeros(0, np.uint32)   # stored visibility (ResultBitSet::m_results)
        self.result_n = 0

    def add(self, lower4, upper4, tidx):
        ext = self.p.box_extent(np.ascontiguousarray(lower4), np.ascontiguousarray(upper4))
        self.lower4 = np.concatenate([self.lower4, lower4]).astype(np.float32)
        self.extent4 = np.concatenate([self.extent4, ext]).astype(np.float32)
        self.tidx = np.concatenate([self.tidx, tidx]).astype(np.uint32)

    def remove(self, index):
        # GroupBitSet::removeObject (GroupBitSet.cpp:93-119): last object moves into the hole,
        # then every attached result moves that object's bit (ResultBitSet.cpp:110-128)
        last = len(self.tidx) - 1
        self.lower4[index] = self.lower4[last]
        self.extent4[index] = self.extent4[last]
        self.tidx[index] = self.tidx[last]
        self.lower4, self.extent4, self.tidx = self.lower4[:last], self.extent4[:last], self.tidx[:last]
        self.p.result_move_bit(self.result, self.result_n, last, index)

    def set_matrices(self, raw, stride=64, count=None):
        self.mats, self.stride = raw, stride

    def matrices_changed(self, indices):
        pass                                    # the port has no OBB cache to invalidate

    def cull(self, vp):
        n = len(self.tidx)
        lower4 = np.ascontiguousarray(self.lower4)
        extent4 = np.ascontiguousarray(self.extent4)
        tidx = np.ascontiguousarray(self.tidx)
        words = self.p.cull_bits(lower4, extent4, tidx, self.mats, vp, self.stride)
        if n != self.result_n:
            self.result = self.p.result_resize(self.result, self.result_n, n)
            self.result_n = n
        changed = self.p.update_changed(words, self.result, n)
        return words, changed

    def bounding_box(self):
        return self.p.bounding_box(np.ascontiguousarray(self.lower4), np.ascontiguousarray(self.extent4),
                                   np.ascontiguousarray(self.tidx), self.mats, self.stride)

    def count(self):
        return len(self.tidx)

    def close(self):
        pass


class CudaEngine:
    kind = "cuda"

    def __init__(self, device=0):
        from pipeline_b200 import capi
        self.capi = capi
        self.ctx = capi.Cull(device)
        self.res = self.ctx.result_create()
        self.lower4 = np.zeros((0, 4), np.float32)
        self.extent4 = np.zeros((0, 4), np.float32)
        self.tidx = np.zeros(0, np.uint32)
        self.dirty = True

    def add(self, lower4, upper4, tidx):
        ext = np.zeros_like(lower4)
        ext[:, :3] = upper4[:, :3] - lower4[:, :3]     # Box3f::getSize in f32 (a6), done by the host layer
        self.lower4 = np.concatenate([self.lower4, lower4]).astype(np.float32)
        self.extent4 = np.concatenate([self.extent4, ext]).astype(np.float32)
        self.tidx = np.concatenate([self.tidx, tidx]).astype(np.uint32)
        self.dirty = True

    def remove(self, index):
        last = len(self.tidx) - 1
        self.lower4[index] = self.lower4[last]
        self.extent4[index] = self.extent4[last]
        self.tidx[index] = self.tidx[last]
        self.lower4, self.extent4, self.tidx = self.lower4[:last], self.extent4[:last], self.tidx[:last]
        self.res.move_bit(last, index)
        self.dirty = True

    def set_matrices(self, raw, stride=64, count=None):
        self.ctx.set_matrices(raw, stride, count)

    def matrices_changed(self, indices):
        self.ctx.update_matrices_from_source(indices)

    def cull(self, vp):
        if self.dirty:
            self.ctx.set_objects(np.ascontiguousarray(self.lower4), np.ascontiguousarray(self.extent4),
                                 np.ascontiguousarray(self.tidx))
            self.dirty = False
        self.ctx.run([self.res], np.ascontiguousarray(vp, np.float32).reshape(1, 16))
        return self.res.bits(), self.res.changed()

    def bounding_box(self):
        if self.dirty:
            self.ctx.set_objects(np.ascontiguousarray(self.lower4), np.ascontiguousarray(self.extent4),
                                 np.ascontiguousarray(self.tidx))
            self.dirty = False
        return self.ctx.bounding_box()

    def count(self):
        return len(self.tidx)

    def close(self):
        self.res.close()
        self.ctx.close()


def run_lifecycle(engine, script, lower4, upper4, mats, frames):
    """Apply tests.cases.lifecycle_script() to an engine; objects use tidx = creation id."""
    out = []
    engine.set_matrices(mats.reshape(-1), 64)
    for step in script:
        if step[0] == "add":
            first, cnt = step[1], step[2]
            engine.add(lower4[first:first + cnt], upper4[first:first + cnt],
                       np.arange(first, first + cnt, dtype=np.uint32))
        elif step[0] == "remove":
            for gi in step[1]:
                engine.remove(gi)
        elif step[0] == "cull":
            bits, changed = engine.cull(frames[step[1]])
            out.append((bits.copy(), changed.copy(), engine.count()))
    return out
