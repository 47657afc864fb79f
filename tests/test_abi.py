"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dpcu.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dpcu.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpcu[A-Z]\w*)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    import __graft_entry__ as g
    g.build_product()
    from pipeline_b200 import capi
    assert os.path.exists(capi.LIB_PATH)
    return capi.LIB_PATH


def test_header_declares_the_survey_abi():
    syms = declared_symbols()
    for needed in ["dpcuDeviceCount", "dpcuBufferCreate", "dpcuHostBufferCreate", "dpcuStreamCreate", "dpcuEventElapsedMs",
                   "dpcuCullCreate", "dpcuCullSetObjects", "dpcuCullSetMatrices", "dpcuCullRun", "dpcuCullResultGetBits",
                   "dpcuCullResultGetChanged", "dpcuCullResultMoveBit", "dpcuTreeCreate", "dpcuTreeSetTopology",
                   "dpcuTreeUpdateLocals", "dpcuTreeCompute", "dpcuTreeWorldDevicePointer", "dpcuCullResultSetPeerBits"]:
        assert needed in syms
    assert len(syms) >= 60


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.check_output(["nm", "-D", "--defined-only", libpath]).decode()
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    lib = C.CDLL(libpath)
    for s in declared_symbols():
        getattr(lib, s)


def test_header_is_plain_c(tmp_path):
    """The boundary must be consumable from C (cgo / JNI / ctypes style FFI): no C++ in the header."""
    src = tmp_path / "t.c"
    src.write_text('#include "dpcu.h"\nint main(void){ dpcuCull *c = 0; (void)c; return dpcuGetVersion() == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_python_binding_covers_the_header(libpath):
    from pipeline_b200 import capi
    L = capi.lib()
    for s in declared_symbols():
        f = getattr(L, s)
        if s not in ("dpcuGetLastError", "dpcuGetVersion"):
            assert f.argtypes is not None, s


def test_no_gpu_fails_loudly(libpath):
    """On a machine without a usable GPU every entry point reports DPCU_ERR_NO_DEVICE - nothing
    silently computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pipeline_b200 import capi
    with pytest.raises(capi.DpcuError) as e:
        capi.device_count()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.DpcuError) as e:
        capi.Cull(0)
    assert e.value.code == -2
    with pytest.raises(capi.DpcuError):
        capi.Tree(0)
    with pytest.raises(capi.DpcuError):
        capi.Buffer(1024)


def test_product_does_not_link_the_oracle(libpath):
    out = subprocess.check_output(["ldd", libpath]).decode()
    assert "dporacle" not in out and "dpref" not in out
    syms = subprocess.check_output(["nm", "-D", libpath]).decode()
    assert "dporacle_" not in syms and "dpref_" not in syms
    banned = ("oracle.loader", "from oracle", "import oracle", "libdporacle", "libdpref", "dporacle_", "dpref_")
    for root, _, files in os.walk(os.path.join(ROOT, "pipeline_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(root, f), errors="replace").read()
                for b in banned:
                    assert b not in txt, "%s references %s: the product path must not use the oracle" % (f, b)


def test_python_constants_follow_the_header():
    """capi.py's DPCU_CULL_OPT_* / DPCU_KERNEL_* numbers are the header's (the harness must not drift from the ABI)."""
    import re
    from pipeline_b200 import capi
    text = open(HEADER).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(DPCU_[A-Z0-9_]+)\s+(\d+)\b", text)}
    pairs = {"DPCU_CULL_OPT_KERNEL": capi.OPT_KERNEL, "DPCU_CULL_OPT_FMA": capi.OPT_FMA, "DPCU_CULL_OPT_CHANGED_LIST": capi.OPT_CHANGED_LIST,
             "DPCU_CULL_OPT_CTAS_PER_SM": capi.OPT_CTAS_PER_SM, "DPCU_CULL_OPT_PROFILE": capi.OPT_PROFILE,
             "DPCU_CULL_OPT_FUSE_LEAF": capi.OPT_FUSE_LEAF, "DPCU_CULL_OPT_FUSE_LIST": capi.OPT_FUSE_LIST,
             "DPCU_CULL_OPT_LAST_KERNEL": capi.OPT_LAST_KERNEL, "DPCU_CULL_OPT_FILTER": capi.OPT_FILTER,
             "DPCU_CULL_OPT_LINE_WORDS": capi.OPT_LINE_WORDS, "DPCU_CULL_OPT_LIST_OFFSETS": capi.OPT_LIST_OFFSETS,
             "DPCU_CULL_OPT_L2_PREFETCH": capi.OPT_L2_PREFETCH,
             "DPCU_KERNEL_AUTO": capi.KERNEL_AUTO, "DPCU_KERNEL_DIRECT": capi.KERNEL_DIRECT, "DPCU_KERNEL_STAGED": capi.KERNEL_STAGED,
             "DPCU_KERNEL_VIEWS": capi.KERNEL_VIEWS, "DPCU_KERNEL_LINES": capi.KERNEL_LINES,
             "DPCU_KERNEL_VIEWS_CHAINS": capi.KERNEL_VIEWS_CHAINS, "DPCU_KERNEL_FUSED_LEAF": capi.KERNEL_FUSED_LEAF,
             "DPCU_KERNEL_LINES_PAIRS": capi.KERNEL_LINES_PAIRS, "DPCU_KERNEL_GRID": capi.KERNEL_GRID}
    for name, value in pairs.items():
        assert defs.get(name) == value, (name, defs.get(name), value)
    options = sorted(v for k, v in defs.items() if k.startswith("DPCU_CULL_OPT_"))
    assert options == list(range(1, len(options) + 1)), "option numbers must be dense and unique: %s" % options
