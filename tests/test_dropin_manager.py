"""The drop-in test: ONE driver (oracle/ref_shim.cpp, the calls dp::sg::xbar::culling::CullingImpl
makes) runs the reference's dp::culling::cpu::Manager and the new dp::culling::cuda::Manager through
the same virtual dp::culling::Manager API; their answers must be identical.

tests/cpp/_build/libdpdropin.so is built in the build container (it compiles the new C++ host layer
against the reference headers where they lie) and travels to the GPU box."""
import os

import numpy as np
import pytest

from pipeline_b200 import scenes
from tests import cases
from tests.engines import RefEngine, run_lifecycle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "tests", "cpp", "_build", "libdpdropin.so")


@pytest.fixture(scope="module")
def dropin():
    from oracle.loader import Reference
    if os.path.isdir("/root/reference"):
        import subprocess
        import __graft_entry__ as g
        g.build_product()
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    if not os.path.exists(DROPIN):
        pytest.skip("tests/cpp/_build/libdpdropin.so not present (built only where /root/reference exists)")
    return Reference(DROPIN)


def test_dropin_library_fails_loudly_without_gpu(dropin):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cpu = dropin.cull(0)           # the reference backend works anywhere
    cpu.close()
    with pytest.raises(RuntimeError) as e:
        dropin.cull(1)             # dp::culling::cuda::Manager::create() throws: no silent CPU path
    assert "no CPU fallback" in str(e.value)
    host_tree = dropin.tree(0)     # likewise dp::transform::Tree vs dp::transform::cuda::Tree
    host_tree.close()
    with pytest.raises(RuntimeError) as e:
        dropin.tree(1)
    assert "no CPU fallback" in str(e.value)


@pytest.mark.gpu
def test_same_driver_two_backends_random_frames(dropin):
    lower4, extent4, upper4, mats, tidx = cases.random_case(50000)
    a, b = RefEngine(dropin, 0), RefEngine(dropin, 1)
    for e in (a, b):
        e.add(lower4, upper4, tidx)
        e.set_matrices(mats.reshape(-1))
    for vp in cases.frames(5):
        ab, ac = a.cull(vp)
        bb, bc = b.cull(vp)
        assert np.array_equal(ab, bb)
        assert np.array_equal(ac, bc)
    assert np.array_equal(a.bounding_box(), b.bounding_box())
    a.close(), b.close()


@pytest.mark.gpu
def test_same_driver_two_backends_lifecycle(dropin, golden):
    g = golden("lifecycle")
    lower4, extent4, upper4, mats, tidx = cases.random_case(5000, seed=0x11FE)
    res = run_lifecycle(RefEngine(dropin, 1), cases.lifecycle_script(), lower4, upper4, mats, cases.frames())
    for k, (bits, changed, n) in enumerate(res):
        assert n == int(g["count%d" % k])
        assert np.array_equal(bits, g["bits%d" % k]), "step %d bits" % k
        assert np.array_equal(changed, g["changed%d" % k]), "step %d changed" % k


@pytest.mark.gpu
def test_same_driver_two_backends_special_and_gather(dropin, golden):
    g = golden("special4k")
    lower4, extent4, upper4, mats, tidx, vps = cases.special_case()
    e = RefEngine(dropin, 1)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    for k, vp in enumerate(vps):
        bits, changed = e.cull(vp)
        assert np.array_equal(bits, g["bits%d" % k])
        assert np.array_equal(changed, g["changed%d" % k])
    e.close()
    g = golden("gather5k")
    lower4, extent4, upper4, raw, tidx, stride = cases.gather_case()
    e = RefEngine(dropin, 1)
    e.add(lower4, upper4, tidx)
    e.set_matrices(raw.reshape(-1), stride, len(raw))
    bits, changed = e.cull(scenes.camera_c2())
    assert np.array_equal(bits, g["bits"]) and np.array_equal(changed, g["changed"])
    assert np.array_equal(e.bounding_box(), g["bbox"])
    e.close()


@pytest.mark.gpu
def test_same_driver_two_backends_growing_group(dropin):
    """The group grows past the capacity of the result's pinned host mirror several times (ResultCUDA::prepare
    re-pins and re-registers it) and shrinks again; bitsets, changed lists and per-object visibility queries
    stay identical to the reference backend."""
    n = 40000
    lower4, extent4, upper4, mats, tidx = cases.random_case(n)
    a, b = RefEngine(dropin, 0), RefEngine(dropin, 1)
    for e in (a, b):
        e.set_matrices(mats.reshape(-1))
    have = 0
    frames = cases.frames(6)
    for k, grow in enumerate((700, 1500, 9000, 28800)):
        for e in (a, b):
            e.add(lower4[have:have + grow], upper4[have:have + grow], tidx[have:have + grow])
        have += grow
        for vp in frames[k:k + 2]:
            ab, ac = a.cull(vp)
            bb, bc = b.cull(vp)
            assert np.array_equal(ab, bb), (k, "bits")
            assert np.array_equal(ac, bc), (k, "changed list")
    for e in (a, b):
        for gi in (0, 39999 - 1, 123, 20000):
            e.remove(gi)
    ab, ac = a.cull(frames[5])
    bb, bc = b.cull(frames[5])
    assert np.array_equal(ab, bb) and np.array_equal(ac, bc)
    a.close(), b.close()


@pytest.mark.gpu
def test_dirty_matrix_protocol(dropin):
    """groupMatrixChanged -> only the flagged matrices are re-read (GroupBitSet.h:140-150)."""
    lower4, extent4, upper4, mats, tidx = cases.random_case(4096)
    mats = mats.copy()
    a, b = RefEngine(dropin, 0), RefEngine(dropin, 1)
    for e in (a, b):
        e.add(lower4, upper4, tidx)
        e.set_matrices(mats.reshape(-1))
    vp = scenes.camera_c2()
    a.cull(vp), b.cull(vp)
    idx = np.arange(0, 4096, 5, dtype=np.uint32)
    mats[idx.astype(np.int64), 3, 2] -= np.float32(250.0)
    for e in (a, b):
        e.set_matrices(mats.reshape(-1))        # same pointer / count / stride: nothing is re-read by itself
        e.matrices_changed(idx)
    ab, ac = a.cull(vp)
    bb, bc = b.cull(vp)
    assert np.array_equal(ab, bb) and np.array_equal(ac, bc) and len(bc) > 0
    a.close(), b.close()


@pytest.mark.gpu
def test_live_object_edit_reaches_the_device(dropin, port):
    """objectSetBoundingBox / objectSetTransformIndex on an object that is already in the group."""
    lower4, extent4, upper4, mats, tidx = cases.random_case(2000)
    e = RefEngine(dropin, 1)
    e.add(lower4, upper4, tidx)
    e.set_matrices(mats.reshape(-1))
    vp = scenes.camera_c2()
    e.cull(vp)
    lower4, upper4, tidx = lower4.copy(), upper4.copy(), tidx.copy()
    for i in (3, 777, 1999):
        lower4[i, :3] -= 50.0
        upper4[i, :3] += 50.0
        tidx[i] = (i * 7) % 2000
        e.set_object(i, lower4[i], upper4[i], int(tidx[i]))
    bits, _ = e.cull(vp)
    ext = port.box_extent(np.ascontiguousarray(lower4), np.ascontiguousarray(upper4))
    want = port.cull_bits(np.ascontiguousarray(lower4), ext, tidx, mats.reshape(-1), vp)
    assert np.array_equal(bits, want)
    e.close()


@pytest.mark.gpu
def test_ten_thousand_live_edits_per_frame(dropin):
    """a frame of an xbar-scale scene edits thousands of live objects.  262 144 objects; every frame
    10 000 objectSetBoundingBox / objectSetTransformIndex calls on live objects, 2 000 removals (swap-remove) and 2 000
    additions, then a cull - through the reference cpu Manager and the cuda Manager, identical bitsets and changed lists
    frame by frame.  The cuda backend ships each frame's edits as ONE batch (dpcuCullSetObjectCount +
    dpcuCullUpdateObjects) and the bit moves of the removals as one batch of words (dpcuCullResultUpdateWords)."""
    import time
    n, extra = 1 << 18, 16000
    lower4, extent4, upper4, mats, tidx = cases.random_case(n + extra, seed=scenes.SEED_C2 + 3)
    a, b = RefEngine(dropin, 0), RefEngine(dropin, 1)
    for e in (a, b):
        e.add(lower4[:n], upper4[:n], tidx[:n])
        e.set_matrices(mats.reshape(-1))
    frames = cases.frames(8)
    for e in (a, b):
        e.cull(frames[0])
    rng = np.random.RandomState(77)
    have, spare = n, n
    spent = {0: 0.0, 1: 0.0}
    culls = {0: 0.0, 1: 0.0}
    for f in range(1, 8):                                   # frames 1 - 3 warm the staging buffers up, 4 - 7 are timed
        edit = rng.choice(have, size=10000, replace=False).astype(np.uint32)
        new_lo = lower4[edit, :3] - rng.uniform(0.0, 30.0, size=(10000, 3)).astype(np.float32)
        new_hi = upper4[edit, :3] + rng.uniform(0.0, 30.0, size=(10000, 3)).astype(np.float32)
        new_ti = rng.randint(0, n + extra, size=10000).astype(np.uint32)
        gone = rng.randint(0, have - 2000, size=2000).astype(np.uint32)      # indices stay valid while the group shrinks
        for k, e in enumerate((a, b)):
            t0 = time.perf_counter()
            e.s.set_objects_many(edit, new_lo, new_hi, new_ti)
            # (the removals / additions that follow dirty the cpu backend's OBB cache; live edits alone would not -
            # the reference quirk of SURVEY.md section 7, hard part 5)
            e.s.remove_objects_many(gone)
            e.add(lower4[spare:spare + 2000], upper4[spare:spare + 2000], tidx[spare:spare + 2000])
            t1 = time.perf_counter()
            e.s.cull(e.r, frames[f])                                # Manager::cull: returns when the result is valid
            if f >= 4:
                spent[k] += time.perf_counter() - t0
                culls[k] += time.perf_counter() - t1
            bits, changed = e.s.visible_bits(e.r), e.s.changed(e.r)  # (the driver's per-object queries: not timed)
            if k == 0:
                want = (bits, changed)
            else:
                assert np.array_equal(bits, want[0]), "frame %d: bits" % f
                assert np.array_equal(changed, want[1]), "frame %d: changed list" % f
        spare += 2000
    print("10 000 edits + 2 000 removals + 2 000 additions + cull per frame over %d objects: reference cpu Manager %.2f ms, "
          "cuda Manager %.2f ms per frame (host wall clock of the 14 000 Manager edit calls + Manager::cull); Manager::cull alone "
          "(cuda: batched upload of the touched objects and words, kernel, result in the host mirror): %.2f ms vs %.2f ms"
          % (n, 250.0 * spent[0], 250.0 * spent[1], 250.0 * culls[0], 250.0 * culls[1]))
    a.close(), b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_interleaved_edits_small_group(dropin, seed):
    """Fuzz of the batched host layer: random interleavings of add / remove / live edit (also edit-then-remove,
    remove-the-last, remove-everything-and-refill) between culls on a small group, 40 frames; cpu Manager and cuda
    Manager must agree on bits, changed list and count after every frame."""
    rng = np.random.RandomState(1000 + seed)
    pool = 3000
    lower4, extent4, upper4, mats, tidx = cases.random_case(pool, seed=scenes.SEED_C2 + 20 + seed)
    tidx = (tidx % np.uint32(pool)).astype(np.uint32)
    a, b = RefEngine(dropin, 0), RefEngine(dropin, 1)
    for e in (a, b):
        e.add(lower4[:400], upper4[:400], tidx[:400])
        e.set_matrices(mats.reshape(-1))
    have, nxt = 400, 400
    frames = cases.frames(8)
    for f in range(40):
        ops = []
        for _ in range(rng.randint(0, 25)):
            kind = rng.choice(["add", "remove", "edit", "edit_remove", "remove_last"], p=[0.3, 0.25, 0.3, 0.1, 0.05])
            if kind == "add" and nxt < pool:
                ops.append(("add", nxt))
                nxt += 1
                have += 1
            elif kind in ("remove", "remove_last", "edit_remove") and have > 0:
                gi = have - 1 if kind == "remove_last" else int(rng.randint(0, have))
                if kind == "edit_remove":
                    ops.append(("edit", gi, int(rng.randint(0, pool)), float(rng.uniform(0.0, 40.0))))
                ops.append(("remove", gi))
                have -= 1
            elif kind == "edit" and have > 0:
                ops.append(("edit", int(rng.randint(0, have)), int(rng.randint(0, pool)), float(rng.uniform(0.0, 40.0))))
        if f == 20:                                             # empty the group, cull it empty, refill
            ops += [("remove", 0)] * have
            have = 0
        if f == 21:
            for _ in range(150):
                if nxt < pool:
                    ops.append(("add", nxt))
                    nxt += 1
                    have += 1
        for e in (a, b):
            for op in ops:
                if op[0] == "add":
                    k = op[1]
                    e.add(lower4[k:k + 1], upper4[k:k + 1], tidx[k:k + 1])
                elif op[0] == "remove":
                    e.remove(op[1])
                else:
                    _, gi, ti, grow = op
                    src = (gi * 7 + ti) % pool
                    e.set_object(gi, lower4[src] - np.float32(grow), upper4[src] + np.float32(grow), ti)
            if any(op[0] == "edit" for op in ops) and e is a:
                e.s.matrices_changed(np.arange(0, 1, dtype=np.uint32))     # reference quirk: only this dirties the cpu OBB cache after a live edit
        ab, ac = a.cull(frames[f % 8])
        bb, bc = b.cull(frames[f % 8])
        assert a.count() == b.count() == have, f
        assert np.array_equal(ab, bb), "frame %d: bits (%d objects, %d ops)" % (f, have, len(ops))
        assert np.array_equal(ac, bc), "frame %d: changed list" % f
    a.close(), b.close()


@pytest.mark.gpu
def test_multi_view_and_device_matrices_extensions(dropin, port):
    from pipeline_b200 import capi
    n = 20000
    lower4, extent4, upper4, mats, tidx = cases.random_case(n, seed=scenes.SEED_C4)
    s = dropin.cull(1)
    s.add_objects(np.ascontiguousarray(lower4[:, :3]), np.ascontiguousarray(upper4[:, :3]), tidx)
    buf = capi.Buffer(mats.nbytes)
    buf.upload(mats)
    s.set_device_matrices(buf.ptr, n)
    results = [s.result_create() for _ in range(6)]
    vps = scenes.cube_map_cameras()
    s.cull_multi(results, vps)
    for v in range(6):
        want = port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), vps[v])
        assert np.array_equal(s.visible_bits(results[v]), want), v
    s.close()
    buf.close()


# ------------------------------------------------------------------ dp::transform::cuda::Tree
def _drive_trees(trees, frames_check, used=None):
    """the same edits on every tree; frames_check(frame) runs after each compute.  used[0] = 1 + the largest
    index handed out so far (slots beyond it are uninitialised storage in the reference)."""
    rng = np.random.RandomState(17)
    used = used if used is not None else [0]

    def locals_for(count, frame):
        return scenes.hierarchy_locals(scenes.SEED_C3 + frame, 1, count, frame=frame)

    # frame 0: build three levels (12 / 120 / 3000 nodes)
    idx = []
    for t in trees:
        l0 = t.add_many(np.zeros(12, np.uint32), locals_for(12, 0))
        l1 = t.add_many(np.repeat(l0, 10), locals_for(120, 1))
        l2 = t.add_many(np.repeat(l1, 25), locals_for(3000, 2))
        idx.append((l0, l1, l2))
        t.compute()
    for a in idx[1:]:
        for x, y in zip(idx[0], a):
            assert np.array_equal(x, y)                  # same index allocation
    l0, l1, l2 = idx[0]
    used[0] = int(l2.max()) + 1
    frames_check(0)
    # frame 1: scattered local edits on every level
    pick = np.unique(np.concatenate([rng.choice(l0, 3), rng.choice(l1, 9), rng.choice(l2, 200)])).astype(np.uint32)
    new = scenes.hierarchy_locals(scenes.SEED_C3 + 5, 1, len(pick), frame=4)
    for t in trees:
        t.update_locals(pick, new)
        t.compute()
    frames_check(1)
    # frame 2: nothing dirty
    for t in trees:
        t.compute()
    frames_check(2)
    # frame 3: topology edits between computes - remove leaves (orphaning nothing), add new subtrees
    gone = rng.choice(l2, 40, replace=False).astype(np.uint32)
    for t in trees:
        for g in gone:
            t.remove(int(g))
        fresh = t.add_many(np.resize(l1, 60), locals_for(60, 6))
        deeper = t.add_many(np.repeat(fresh[:10], 4), locals_for(40, 7))        # a fourth level
        t.update_locals(l0[:2], locals_for(2, 8))
        t.compute()
        used[0] = max(used[0], int(fresh.max()) + 1, int(deeper.max()) + 1)
    frames_check(3)
    # frame 4: a contiguous run of locals (whole level) rewritten
    for t in trees:
        t.update_locals(l1, locals_for(len(l1), 9))
        t.compute()
    frames_check(4)
    # frame 5: remove + add that leaves every level size unchanged (the case only the explicit
    # topology flag catches when edits go through the derived type)
    victim = int(l2[7])
    for t in trees:
        t.remove(victim)
        again = t.add_many(np.asarray([l1[3]], np.uint32), locals_for(1, 10))
        t.compute()
    frames_check(5)


@pytest.mark.gpu
@pytest.mark.parametrize("backend", [1, 2])
def test_transform_tree_dropin_same_driver(dropin, backend):
    """dp::transform::Tree and dp::transform::cuda::Tree behind the same calls: world matrices, the
    published dirty set (EventWorldMatricesChanged) and index allocation must be identical frame by
    frame - for edits through the derived type (backend 1) and through a base reference (backend 2;
    the final same-size remove + add is skipped there, see Tree.h)."""
    ref, dev = dropin.tree(0), dropin.tree(backend)
    used = [0]

    def check(frame):
        if backend == 2 and frame == 5:
            return
        assert ref.count() == dev.count()
        assert np.array_equal(ref.dirty_world(), dev.dirty_world()), (frame, "published dirty set")
        assert np.array_equal(ref.world()[:used[0]].view(np.uint32), dev.world()[:used[0]].view(np.uint32)), (frame, "world matrices")

    _drive_trees([ref, dev], check, used)
    ref.close(), dev.close()


@pytest.mark.gpu
def test_transform_tree_feeds_cuda_manager_zero_copy(dropin):
    """SURVEY.md 8f rank 1 + 3: the reference frame loop (TransformTree::compute, then CullingImpl::cull with
    the tree's world matrices) on both stacks - cpu: Tree -> groupSetMatrices(host world) -> cpu::Manager;
    cuda: cuda::Tree -> groupSetDeviceMatrices(device world) -> cuda::Manager, no matrix ever re-uploaded,
    host world mirror off.  Bitsets and changed lists must match every frame."""
    ref_tree, dev_tree = dropin.tree(0), dropin.tree(1)
    dev_tree.host_mirror(False)
    cpu, gpu = dropin.cull(0), dropin.cull(1)
    rc, rg = cpu.result_create(), gpu.result_create()
    state = {"leaves": None}
    view = scenes.make_look_at((0, 0, 30), (0, 0, 0), (0, 1, 0))      # inside the cloud of nodes: some are behind the eye
    vps = [scenes.mat_mul(view, scenes.make_perspective(fov, 1.3, 1.0, 600.0)) for fov in (35.0, 15.0, 50.0, 35.0, 25.0, 45.0)]

    def check(frame):
        n_nodes = ref_tree.count()
        if frame == 0:
            # one object per node of the deepest level that exists after the first frame
            rng = np.random.RandomState(2)
            n = 3000
            lower4, extent4, upper4, _, _ = cases.random_case(n)
            tidx = rng.randint(1, 1 + 12 + 120 + 3000, size=n).astype(np.uint32)
            for s in (cpu, gpu):
                s.add_objects(np.ascontiguousarray(lower4[:, :3] * 0.3), np.ascontiguousarray(upper4[:, :3] * 0.3), tidx)
        cpu.set_matrices(ref_tree.world_view(), 64, n_nodes)
        for i in np.flatnonzero(np.unpackbits(ref_tree.dirty_world().view(np.uint8), bitorder="little")):
            cpu.matrix_changed(int(i))                                   # what the dormant TransformObserver would do
        gpu.set_device_matrices(dev_tree.device_world(), n_nodes)
        cpu.cull(rc, vps[frame])
        gpu.cull(rg, vps[frame])
        assert np.array_equal(cpu.visible_bits(rc), gpu.visible_bits(rg)), frame
        assert np.array_equal(cpu.changed(rc), gpu.changed(rg)), frame
        if frame == 0:
            vis = int(np.unpackbits(cpu.visible_bits(rc).view(np.uint8)).sum())
            assert 0 < vis < 3000

    _drive_trees([ref_tree, dev_tree], check)
    cpu.close(), gpu.close(), ref_tree.close(), dev_tree.close()


# ------------------------------------------------------------------ the same, as the reference's own C++ would write it
FRAME_LOOP = os.path.join(ROOT, "tests", "cpp", "_build", "frame_loop")


def test_cpp_frame_loop_fails_loudly_without_gpu(dropin):
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(FRAME_LOOP):
        pytest.skip("tests/cpp/_build/frame_loop not present")
    r = subprocess.run([FRAME_LOOP], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_frame_loop_reference_stack_vs_cuda_stack(dropin):
    """tests/cpp/frame_loop.cpp: dp::transform::Tree + dp::culling::cpu::Manager next to dp::transform::cuda::Tree +
    dp::culling::cuda::Manager, driven by plain C++ exactly as xbar drives them; six frames, everything identical."""
    import subprocess
    if not os.path.exists(FRAME_LOOP):
        pytest.skip("tests/cpp/_build/frame_loop not present (built only where /root/reference exists)")
    r = subprocess.run([FRAME_LOOP], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok") and r.stdout.count("identical") == 6, r.stdout


# ------------------------------------------------------------------ through the PATCHED dp::sg::xbar::culling::CullingImpl
XBAR_LOOP = os.path.join(ROOT, "tests", "cpp", "_build", "xbar_loop")


def test_patched_culling_impl_fails_loudly_without_gpu(dropin):
    """Mode::CUDA in the patched CullingImpl has no silent CPU path: without a device Culling::create throws."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(XBAR_LOOP):
        pytest.skip("tests/cpp/_build/xbar_loop not present")
    r = subprocess.run([XBAR_LOOP], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_patched_culling_impl_cpu_mode_vs_cuda_mode(dropin):
    """SURVEY.md 8 a15 / f1 as code: the reference's own CullingImpl.cpp with patches/0001-culling-cuda-backend.patch
    applied (case Mode::CUDA, device-resident matrix feed, TransformObserver attached), compiled against stand-in
    SceneTree headers and driven through Culling::create / cull / resultGetChangedIndices / resultIsVisible /
    getBoundingBox with SceneTree ADDED / REMOVED / CHANGED events: Mode::CPU and Mode::CUDA agree on eight frames."""
    import subprocess
    if not os.path.exists(XBAR_LOOP):
        pytest.skip("tests/cpp/_build/xbar_loop not present (built only where /root/reference exists)")
    r = subprocess.run([XBAR_LOOP], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok") and r.stdout.count("identical") == 8, r.stdout
