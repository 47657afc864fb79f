"""Static checks on the machine code of the built library (no GPU needed): the exact-mode cull
kernels must not contain a contracted multiply-add, and the staged kernel must really use the
TMA bulk-copy / mbarrier / cp.async path it claims.

Why this exists: ptxas (CUDA 12.9) fuses `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 even under
--fmad=false, which rounds once instead of twice and breaks bit-exactness against the reference.
pipeline_b200/csrc/cull_views.cuh::addProd keeps every add-of-a-product as fma(x, one, y) with a
runtime `one`; this test pins that property on the SASS."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pipeline_b200", "lib", "libdpcu.so")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or libdpcu.so not available")
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
            ADDRS[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line.split("*/", 1)[1].split("/*")[0].strip())
            ADDRS[name].append(int(re.match(r"\s+/\*([0-9a-f]{4})\*/", line).group(1), 16))
    return funcs


ADDRS = {}            # function name -> address of every instruction of its body (parallel to the body list)


def _argument_source(name, body, pos, reg):
    """`reg` is read at `pos` inside a local callee (the code behind a CALL.REL target, e.g. the queued exact pass
    mvFlush of the multi-view lines kernel) and is not written between the callee's entry and `pos`: it is an incoming
    argument.  Returns the constant-bank offsets it was loaded from before every call site (None if any call site
    defines it differently)."""
    addrs = ADDRS[name]
    calls = [(k, int(m.group(1), 16)) for k, i in enumerate(body) for m in [re.search(r"CALL\.REL\.NOINC\s+(0x[0-9a-f]+)", i)] if m]
    targets = sorted({t for _, t in calls if t <= addrs[pos]})
    if not targets:
        return None
    entry = addrs.index(targets[-1])
    for back in range(pos - 1, entry - 1, -1):
        d = _DEST.match(body[back])
        if d and d.group(1) == reg:
            return None
    sources = []
    for k, t in calls:
        if t != targets[-1]:
            continue
        src = _constant_source(body, k, reg)
        if src is None:
            return None
        sources.append(src)
    return sources


def _kernels(funcs, stem):
    return {k: v for k, v in funcs.items() if stem in k}


PARAM_BASE = 0x380            # kernel parameters start here in constant bank 0 on sm_100


def _arg_layout(nv):
    import ctypes as C
    L = C.CDLL(LIB)
    one, vp, flt, total = (C.c_size_t() for _ in range(4))
    L.dpcuDebugKernelArgLayout.argtypes = [C.c_int] + [C.POINTER(C.c_size_t)] * 4
    assert L.dpcuDebugKernelArgLayout(nv, C.byref(one), C.byref(vp), C.byref(flt), C.byref(total)) == 0
    return {"one": PARAM_BASE + one.value, "vp": PARAM_BASE + vp.value, "filter": PARAM_BASE + flt.value,
            "end": PARAM_BASE + total.value}


_LOAD = re.compile(r"(?:@!?U?P\d+\s+)?(LDCU|LDC)(?:\.(\d+))?\s+(U?R\d+),\s*c\[0x0\]\[(?:(U?R\d+)\+)?(0x[0-9a-f]+)\]")
_DEST = re.compile(r"(?:@!?U?P\d+\s+)?\S+\s+(U?R\d+)\b")


def _constant_source(body, pos, reg):
    """Constant-bank offset the nearest preceding definition of `reg` loaded it from (None: not a constant load;
    a wide load covers the following registers as well).  Linear scan - good enough for straight-line loop bodies."""
    kind, num = ("UR", int(reg[2:])) if reg.startswith("UR") else ("R", int(reg[1:]))
    for back in range(pos - 1, -1, -1):
        m = _LOAD.match(body[back])
        if m:
            width = int(m.group(2) or 32) // 32
            dk, dn = ("UR", int(m.group(3)[2:])) if m.group(3).startswith("UR") else ("R", int(m.group(3)[1:]))
            if dk == kind and dn <= num < dn + width:
                return int(m.group(5), 16)
            continue
        d = _DEST.match(body[back])
        if d and d.group(1) == reg:
            return None
    return None


def test_exact_kernels_have_no_contracted_multiply_add(sass):
    """The reference arithmetic must round every product and every sum separately.  In the kernels that carry
    it, a fused multiply-add may only be (a) the packed add-of-a-product idiom: FFMA2 (packed product) * one +
    (packed product) with `one` = CullArgs::onePair, or (b) arithmetic of the multi-view filter (cull_filter.cuh),
    recognisable by an operand loaded from the filter constants; in particular nothing that multiplies by a
    view-projection entry may be fused."""
    exact = {k: v for k, v in sass.items()
             if any(s in k for s in ("cullDirectKernel", "cullViewsKernel", "cullStagedKernel", "cullLinesKernel",
                                     "cullLinesMvKernel", "mvFlush", "cullFusedLeafKernel", "treeLevelKernel", "treeLevelWideKernel",
                                     "boundingBoxKernel")) and "_fma" not in k}
    assert len(exact) >= 8 * 7 + 3
    checked_filter = 0
    for name, body in exact.items():
        m = re.search(r"Kernel(?:I|_fmaI)Li(\d)E", name)
        layout = _arg_layout(int(m.group(1))) if m else None
        for pos, i in enumerate(body):
            mm = re.match(r"(?:@!?U?P\d+\s+)?(FFMA2?)\b(.*)", i)
            if not mm:
                continue
            packed = mm.group(1) == "FFMA2"
            ops = [o.strip() for o in i.split(None, 2 if i.startswith("@") else 1)[-1].rstrip(" ;").split(",")]
            assert layout is not None, "%s contains %s" % (name, i)
            regs = [re.match(r"-?\|?(U?R\d+)", o).group(1) for o in ops[1:] if re.match(r"-?\|?(U?R\d+)", o)]
            sources = [_constant_source(body, pos, r) for r in regs]
            if any(src is not None and layout["filter"] <= src < layout["end"] for src in sources):
                checked_filter += 1                       # (b) the filter's own arithmetic
                continue
            if not packed and len(regs) >= 2 and regs[0] == regs[1]:
                checked_filter += 1                       # (b') a square (the filter's radius): the reference never squares
                continue
            # (a) addProd
            assert packed, "%s: scalar fused multiply-add outside the filter: %s" % (name, i)
            assert len(ops) == 4 and ops[1].endswith(".F32x2.HI_LO") and ops[1].startswith("R"), "%s: %s" % (name, i)
            if not ops[2].startswith("UR"):
                # the queued exact pass is a local callee: `one` arrives in a register pair, loaded from onePair before each call
                lo = ops[2].split(".")[0]
                hi = "R%d" % (int(lo[1:]) + 1)
                for reg, base in ((lo, layout["one"]), (hi, layout["one"] + 4)):
                    srcs = _argument_source(name, body, pos, reg)
                    assert srcs and all(x == base for x in srcs), "%s: FFMA2 multiplier %s is not CullArgs::onePair: %s" % (name, reg, i)
                continue
            src = _constant_source(body, pos, ops[2].split(".")[0])
            if src is None:                               # a UMOV of a half loaded elsewhere: follow it once
                for back in range(pos - 1, -1, -1):
                    u = re.match(r"(?:@!?U?P\d+\s+)?UMOV\s+%s,\s*(UR\d+)" % ops[2].split(".")[0], body[back])
                    if u:
                        src = _constant_source(body, back, u.group(1))
                        break
            assert src is not None and layout["one"] <= src < layout["one"] + 8, \
                "%s: FFMA2 multiplier is not CullArgs::onePair (%s): %s" % (name, src, i)
    assert checked_filter > 0


def test_fma_variant_is_really_fused(sass):
    fused = _kernels(sass, "cullDirectKernel_fma")
    assert fused
    assert all(any(i.startswith("FFMA") for i in body) for body in fused.values())


def test_views_kernel_uses_packed_arithmetic(sass):
    for name, body in _kernels(sass, "cullViewsKernel").items():
        if "ILi1E" in name:
            continue
        ops = [i.split()[0] if not i.startswith("@") else i.split()[1] for i in body]
        assert ops.count("FMUL2") >= 56 and ops.count("FADD2") >= 30, name


def test_staged_kernel_uses_tma_bulk_copies(sass):
    staged = _kernels(sass, "cullStagedKernel")
    assert len(staged) == 8
    for name, body in staged.items():
        text = "\n".join(body)
        assert "UBLKCP" in text, name + ": no bulk TMA copy (cp.async.bulk)"
        assert "SYNCS" in text, name + ": no mbarrier"
        assert "LDGSTS" in text, name + ": no cp.async gather"


def test_gathered_matrices_are_read_with_256_bit_loads(sass):
    """every gather-by-index matrix read is two LDG.E.256 (ld.global.nc.v8.f32), not four 128-bit row
    loads - half the L1 wavefronts of a warp whose matrices lie 64 bytes apart."""
    for family in ("cullDirectKernel", "cullLinesKernel", "cullLinesMvKernel", "cullViewsKernel", "cullFusedLeafKernel"):
        kernels = {n: b for n, b in _kernels(sass, family).items() if "_fma" not in n}
        assert kernels, family
        for name, body in kernels.items():
            wide = [i for i in body if re.search(r"\bLDG\.E\.[A-Z0-9.]*256", i)]
            assert len(wide) >= 2, name + ": no 256-bit matrix loads"


def test_line_granular_kernels_prefetch_full_sectors(sass):
    """The single-view line-granular kernel and the pair-filter kernel up to four views look ahead: at least three L2
    prefetches per step (two sectors of the next matrix, one of the extents); five and more views keep the plain form."""
    def prefetches(body):
        return sum(1 for i in body if re.search(r"\bCCTL\.[A-Z0-9.]*PF2|\bPREFETCH|CCTL\.E\.PF2", i))
    one = [b for n, b in _kernels(sass, "cullLinesKernel").items() if "ILi1E" in n]
    assert one and all(prefetches(b) >= 3 for b in one)
    for name, body in _kernels(sass, "cullLinesMvKernel").items():
        nv = int(re.search(r"ILi(\d)E", name).group(1))
        if nv <= 4:
            assert prefetches(body) >= 3, name
        else:
            assert prefetches(body) == 0, name
