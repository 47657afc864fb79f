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
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line.split("*/", 1)[1].split("/*")[0].strip())
    return funcs


def _kernels(funcs, stem):
    return {k: v for k, v in funcs.items() if stem in k}


def test_exact_kernels_have_no_contracted_multiply_add(sass):
    exact = {k: v for k, v in sass.items()
             if any(s in k for s in ("cullDirectKernel", "cullViewsKernel", "cullStagedKernel", "treeLevelKernel",
                                     "boundingBoxKernel")) and "_fma" not in k}
    assert len(exact) >= 8 * 3 + 2
    for name, body in exact.items():
        scalar = [i for i in body if re.match(r"(@!?U?P\d+\s+)?FFMA\b", i)]
        assert not scalar, "%s contains scalar FFMA: %s" % (name, scalar[:3])
        # every FFMA2 must be addProd: (packed product) * one + (packed product).  `one` arrives as a kernel
        # argument: the uniform register used as multiplier must have been loaded, by the nearest preceding
        # definition in program order, from ONE fixed constant-bank address per kernel (CullArgs::onePair),
        # never from the indexed view-projection rows, and the multiplicand is never a scalar broadcast
        sources = set()
        for pos, i in enumerate(body):
            if not re.match(r"(@!?U?P\d+\s+)?FFMA2\b", i):
                continue
            ops = [o.strip() for o in i.split(None, 1)[1].rstrip(" ;").split(",")]
            assert len(ops) == 4, i
            assert ops[1].endswith(".F32x2.HI_LO") and ops[1].startswith("R"), "%s: %s" % (name, i)
            assert ops[2].startswith("UR"), "%s: %s" % (name, i)
            ur = ops[2].split(".")[0]
            for back in range(pos - 1, -1, -1):
                m = re.match(r"(@!?U?P\d+\s+)?(LDCU(?:\.\d+)?|UMOV)\s+%s,\s*(\S+)\s*;" % ur, body[back])
                if m:
                    if m.group(2) != "UMOV":            # (a UMOV copies a half loaded elsewhere)
                        sources.add(m.group(3))
                    break
            else:
                raise AssertionError("%s: no definition of %s before %s" % (name, ur, i))
        assert len(sources) <= 1, "%s: FFMA2 multipliers come from %s" % (name, sources)
        assert all("UR" not in a for a in sources), "%s: FFMA2 multiplier loaded from an indexed address %s" % (name, sources)


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
