"""patches/0001-culling-cuda-backend.patch - the reference-side hookup as code: it must apply
cleanly to the reference tree, be what tools/make_patch.py generates, and leave a CullingImpl that selects the cuda
backend.  Needs /root/reference (build container); skipped elsewhere - nothing here runs on the GPU box."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "patches", "0001-culling-cuda-backend.patch")
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None,
                                reason="needs the reference tree and patch(1)")


def _touched():
    files = []
    for line in open(PATCH, "rb").read().decode("latin-1").splitlines():
        if line.startswith("+++ b/"):
            files.append(line[6:].strip())
    return files


def test_patch_applies_cleanly_to_the_reference(tmp_path):
    files = _touched()
    assert "dp/sg/xbar/culling/src/CullingImpl.cpp" in files and "dp/sg/renderer/rix/gl/src/DrawableManagerDefault.cpp" in files
    assert "dp/sg/xbar/TransformTree.h" in files and "dp/culling/CMakeLists.txt" in files
    for f in files:
        dst = tmp_path / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), dst)
    for extra in (["--dry-run"], []):
        r = subprocess.run(["patch", "-p1", "--forward"] + extra + ["-i", PATCH], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0 and "FAILED" not in r.stdout and "fuzz" not in r.stdout, r.stdout + r.stderr
    impl = (tmp_path / "dp/sg/xbar/culling/src/CullingImpl.cpp").read_bytes().decode("latin-1")
    assert "case dp::culling::Mode::CUDA:" in impl and "dp::culling::cuda::Manager::create()" in impl
    assert "getTransformTree().getTree().attach( m_transformObserver.get() )" in impl          # the dormant observer is attached
    assert "groupSetDeviceMatrices" in impl                                                   # device-resident feed
    dm = (tmp_path / "dp/sg/renderer/rix/gl/src/DrawableManagerDefault.cpp").read_bytes().decode("latin-1")
    assert "Culling::create( getSceneTree(), cullingMode )" in dm                             # AUTO passes the resolved mode
    # applying it a second time must be refused (it is not a no-op patch)
    r = subprocess.run(["patch", "-p1", "--forward", "--dry-run", "-i", PATCH], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0


def test_patch_is_what_the_generator_writes(tmp_path, monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_patch
    monkeypatch.setattr(make_patch, "OUT", str(tmp_path / "again.patch"))
    make_patch.main()
    assert open(PATCH, "rb").read() == open(tmp_path / "again.patch", "rb").read()
