"""BASELINE.json's full sizes on the GPU (C3: 17.9 M-node tree + 16 Mi objects, C4: 64 Mi objects x 6
views, C5: 256 Mi objects): the oracle cannot cull these in seconds, so parity is carried by
  * sampled slices - the device-generated scene is replayed on the host for randomly placed
    32 768-object slices (tests/test_scene_gen.py pins generator == replay) and the oracle's words
    for those slices must equal the words the GPU produced at the same positions, bit for bit;
  * size-independent properties - changed list == ascending set bits of (new ^ old), its length ==
    popcount, a repeated cull changes nothing, one six-view pass == six single-view passes, every
    kernel form produces the same words, and a scene culled in shards == the scene culled whole
    (the multi-GPU partitioning of SURVEY.md 8e, on one device)."""
import zlib

import numpy as np
import pytest

from pipeline_b200 import scenes

pytestmark = pytest.mark.gpu

SLICE = 32768


@pytest.fixture(scope="module")
def capi():
    from pipeline_b200 import capi
    assert capi.device_count() >= 1
    return capi


def _set_bits(words, n):
    return np.flatnonzero(np.unpackbits(words.view(np.uint8), bitorder="little")[:n]).astype(np.uint32)


def _popcount(words):
    return int(np.unpackbits(words.view(np.uint8)).sum())


def _slice_starts(n, k, seed):
    """first, last (ragged tail included) and k random 1024-aligned slice starts"""
    rng = np.random.RandomState(seed)
    last = max(((n - 1) // 1024) * 1024 - SLICE + 1024, 0)
    starts = [0, last] + [int(x) * 1024 for x in rng.randint(0, max((n - SLICE) // 1024, 1), size=k)]
    return sorted(set(starts))


class _Scene:
    """n objects of the C2/C4/C5 family generated on the device (object i -> matrix i)"""

    def __init__(self, capi, seed, n, first=0):
        self.capi, self.seed, self.n, self.first = capi, seed, n, first
        self.lo, self.ex, self.mt = capi.Buffer(n * 16), capi.Buffer(n * 16), capi.Buffer(n * 64)
        capi.scene_generate(seed, first, n, first, self.lo.ptr, self.ex.ptr, self.mt.ptr)
        capi.device_sync()
        self.ctx = capi.Cull(0)
        self.ctx.set_objects(self.lo.ptr, self.ex.ptr, None, capi.MEM_DEVICE, n=n)
        self.ctx.bind_matrices(self.mt.ptr, n)

    def oracle_words(self, port, start, vp):
        """the oracle's visibility words of objects [start, start + SLICE) (clipped to n)"""
        count = min(SLICE, self.n - start)
        lower4, extent4, _, mats, _ = scenes.random_objects(self.seed, self.first + start, count)
        tidx = np.arange(count, dtype=np.uint32)
        return port.cull_bits(lower4, extent4, tidx, mats.reshape(-1), vp)

    def oracle_words_contiguous(self, port, start, count, vp_list, threads=8):
        """SURVEY.md 8d parity reporting: one contiguous run of >= 2^24 objects, word for word.  The run's inputs are
        the bytes the GPU culled (downloaded; the sampled slices above pin them to the host replay of the generator),
        culled by the oracle with `threads` host threads."""
        lo = self.lo.download(np.empty((count, 4), np.float32), offset=start * 16)
        ex = self.ex.download(np.empty((count, 4), np.float32), offset=start * 16)
        mt = self.mt.download(np.empty(count * 16, np.float32), offset=start * 64)
        tidx = lo[:, 3].copy().view(np.uint32) - np.uint32(start)
        assert np.array_equal(tidx, np.arange(count, dtype=np.uint32))
        lo[:, 3] = 1.0
        return [port.cull_bits(lo, ex, tidx, mt, vp, threads=threads) for vp in vp_list]

    def close(self):
        self.ctx.close()
        for b in (self.lo, self.ex, self.mt):
            b.close()


def _check_slices(scene, port, words, vp, starts):
    for s in starts:
        want = scene.oracle_words(port, s, vp)
        got = words[s // 32: s // 32 + len(want)]
        assert np.array_equal(got, want), "slice at object %d differs from the oracle" % s


def _check_changed(old_words, new_words, changed, n):
    flips = new_words ^ old_words
    assert len(changed) == _popcount(flips)
    assert np.array_equal(changed, _set_bits(flips, n))          # ascending group indices (BitArray::traverseBits order)


@pytest.mark.parametrize("n", [(1 << 26), (1 << 26) - 12345])
def test_c4_sixty_four_million_single_view(capi, port, n):
    sc = _Scene(capi, scenes.SEED_C4, n)
    r = sc.ctx.result_create()
    words = (n + 31) // 32
    old = np.full(words, 0xFFFFFFFF, np.uint32)                  # new objects start visible (ResultBitSet.cpp:65-79)
    if n & 31:
        old[-1] = (1 << (n & 31)) - 1
    starts = _slice_starts(n, 6, seed=n & 0xFFFF)
    for frame in (0, 5):
        vp = scenes.orbit_camera(frame)
        sc.ctx.run([r], vp)
        new = r.bits()
        assert 0.02 < _popcount(new) / n < 0.5
        if n & 31:
            assert new[-1] >> (n & 31) == 0                      # unused tail bits stay 0 (BitArray.h:298-308)
        _check_slices(sc, port, new, vp, starts)
        _check_changed(old, new, r.changed(), n)
        old = new
    # idempotence: the same camera again changes nothing
    sc.ctx.run([r], scenes.orbit_camera(5))
    assert r.changed_count() == 0
    assert np.array_equal(r.bits(), old)
    # the device-built visible-instance list is the ascending set bits
    r.build_visible_list()
    assert np.array_equal(r.visible(), _set_bits(old, n))
    # every kernel form yields the same words at full size
    crc = zlib.crc32(old.tobytes())
    hidden = _set_bits(~old, n)
    for kernel in (capi.KERNEL_DIRECT, capi.KERNEL_VIEWS, capi.KERNEL_LINES, capi.KERNEL_STAGED):
        sc.ctx.set_option(capi.OPT_KERNEL, kernel)
        r2 = sc.ctx.result_create()
        sc.ctx.run([r2], scenes.orbit_camera(5))
        assert zlib.crc32(r2.bits().tobytes()) == crc, "kernel form %d" % kernel
        assert r2.changed_count() == n - _popcount(old)          # first cull of a result: the invisible set
        if kernel == capi.KERNEL_LINES:
            # the list built inside the kernel (look-back over 65 536 lines): first cull = every hidden object
            # (58 M entries), then a moved camera = the ascending flips
            assert np.array_equal(r2.changed(), hidden)
            sc.ctx.run([r2], scenes.orbit_camera(9))
            _check_changed(old, r2.bits(), r2.changed(), n)
        r2.close()
    r.close()
    sc.close()


def test_c4_sixty_four_million_six_views(capi, port):
    n = 1 << 26
    sc = _Scene(capi, scenes.SEED_C4, n)
    res = [sc.ctx.result_create() for _ in range(6)]
    words = n // 32
    old = [np.full(words, 0xFFFFFFFF, np.uint32) for _ in range(6)]
    starts = _slice_starts(n, 2, seed=66)
    for frame, eye in enumerate(((0.0, 0.0, 0.0), (30.0, -10.0, 5.0))):
        vps = np.ascontiguousarray(scenes.cube_map_cameras(eye), np.float32)
        sc.ctx.run(res, vps)
        new = [r.bits() for r in res]
        for v in range(6):
            _check_slices(sc, port, new[v], vps[v], starts)
            _check_changed(old[v], new[v], res[v].changed(), n)
        if frame == 1:
            # one contiguous 2^24-object run of every view against the oracle, word for word (100 M decisions)
            start, count = 19 * (1 << 20), 1 << 24
            want = sc.oracle_words_contiguous(port, start, count, list(vps))
            for v in range(6):
                bad = np.flatnonzero(new[v][start // 32:(start + count) // 32] != want[v])
                assert bad.size == 0, "view %d: %d words differ in the contiguous run, first at word %d" % (v, bad.size, start // 32 + bad[0])
        if frame == 0:
            # the six 90-degree frusta from one eye cover all directions: every object is in at least one
            # of them unless it lies beyond the far plane (the scene reaches |x| = 1000 * sqrt(3) > 1500)
            union = np.bitwise_or.reduce(np.stack(new))
            assert _popcount(union) / n > 0.95
        old = new
    # the multi-view filter decides only what it can prove: reference arithmetic for every pair (filter off) and
    # the filter with 1/8 of its margin (the rounding-error bound of the proof itself) give the same words
    for mode in (0, 2):
        sc.ctx.set_option(capi.OPT_FILTER, mode)
        again = [sc.ctx.result_create() for _ in range(6)]
        sc.ctx.run(again, vps)
        for v in range(6):
            assert np.array_equal(again[v].bits(), old[v]), (mode, v)
        for r in again:
            r.close()
    sc.ctx.set_option(capi.OPT_FILTER, 1)
    # one six-view pass == six single-view passes (new results: compare words only)
    single = sc.ctx.result_create()
    for v in range(6):
        sc.ctx.run([single], vps[v])
        assert np.array_equal(single.bits(), old[v]), v
    single.close()
    for r in res:
        r.close()
    sc.close()


def test_c5_quarter_billion_sharded_equals_whole(capi, port):
    """C5: 2^28 objects.  The scene culled whole on one GPU == the same scene culled as 4 contiguous
    shards (separate contexts over their own slices, as the multi-GPU bench runs them); sampled
    slices against the oracle on the whole-scene result."""
    n = 1 << 28
    vp = scenes.orbit_camera(2)
    sc = _Scene(capi, scenes.SEED_C5, n)
    r = sc.ctx.result_create()
    sc.ctx.run([r], vp)
    whole = r.bits()
    assert r.changed_count() == n - _popcount(whole)
    _check_slices(sc, port, whole, vp, _slice_starts(n, 4, seed=5))
    start, count = 150 * (1 << 20) + 7168, 1 << 24               # a contiguous 2^24-object run, word for word
    want = sc.oracle_words_contiguous(port, start, count, [vp])[0]
    assert np.array_equal(whole[start // 32:(start + count) // 32], want)
    del want
    changed = r.changed()
    assert np.array_equal(changed, _set_bits(~whole, n))
    r.close()
    sc.close()
    del changed
    shards = 4
    per = n // shards
    for g in range(shards):
        sh = _Scene(capi, scenes.SEED_C5, per, first=g * per)
        rs = sh.ctx.result_create()
        sh.ctx.run([rs], vp)
        assert np.array_equal(rs.bits(), whole[g * per // 32:(g + 1) * per // 32]), "shard %d" % g
        rs.close()
        sh.close()


def test_c3_full_tree_and_sixteen_million_objects(capi, port):
    """C3: 4-level fan-out-16 tree (17 895 424 nodes + root), every local matrix dirty, 16 Mi objects
    bound one per leaf; the last level runs inside the cull kernel.  Sampled runs of 1024 leaves: the
    oracle propagates their ancestor chains and culls them; world matrices and visibility words must
    match bit for bit.  Then a frame that dirties nothing must leave everything as it was."""
    levels = (4096, 65536, 1048576, 16777216)
    entries, offsets, n_nodes = scenes.hierarchy_topology(levels)
    n = levels[-1]
    first_leaf = n_nodes - n
    tree = capi.Tree(0)
    tree.set_topology(entries, offsets, n_nodes)
    lo, ex, scratch = capi.Buffer(n * 16), capi.Buffer(n * 16), capi.Buffer(n * 64)
    capi.scene_generate(scenes.SEED_C3, 0, n, (-first_leaf) & 0xFFFFFFFF, lo.ptr, ex.ptr, scratch.ptr)
    capi.device_sync()
    scratch.close()
    lo2, ex2, locbuf = capi.Buffer(n_nodes * 16), capi.Buffer(n_nodes * 16), capi.Buffer(n_nodes * 64)
    capi.scene_generate(scenes.SEED_C3 + 1, 0, n_nodes, 0, lo2.ptr, ex2.ptr, locbuf.ptr)   # locals: rigid placements, node i <- draw i
    capi.device_sync()
    lo2.close(), ex2.close()
    tree.set_locals(0, locbuf.ptr, capi.MEM_DEVICE, count=n_nodes)
    all_local = locbuf.download(np.empty((n_nodes, 4, 4), np.float32))      # the bytes the GPU propagates, for the oracle below
    locbuf.close()
    tree.mark_dirty(1, n_nodes - 1)
    ctx = capi.Cull(0)
    ctx.set_objects(lo.ptr, ex.ptr, None, capi.MEM_DEVICE, n=n)
    wptr, cnt = tree.world_ptr()
    ctx.bind_matrices(wptr, cnt)
    r = ctx.result_create()
    vp = scenes.orbit_camera(1)
    launches = tree.launches()
    ctx.run_with_tree(tree, [r], vp)
    words = r.bits()
    assert tree.launches() - launches >= 4                       # three level kernels + the fused leaf level
    dirty = tree.dirty_world()
    assert _popcount(dirty) == n_nodes - 1                       # every node but the root was recomputed

    def local_of(nodes):
        out = np.zeros((len(nodes), 4, 4), np.float32)
        for k, node in enumerate(nodes):
            out[k] = scenes.random_objects(scenes.SEED_C3 + 1, int(node), 1)[3][0]
        return out

    rng = np.random.RandomState(33)
    for s in [0, n - 1024] + [int(x) * 1024 for x in rng.randint(0, n // 1024, size=3)]:
        leaves = np.arange(first_leaf + s, first_leaf + s + 1024, dtype=np.uint64)
        # ancestors through the level-sorted entry list (entry e describes node e + 1)
        chain = [leaves]
        for _ in range(3):
            chain.append(np.unique(entries[chain[-1] - 1, 0].astype(np.uint64)))
        nodes = np.concatenate([[0]] + chain[::-1]).astype(np.uint64)      # root, level 0 .. leaves: level order
        remap = {int(g): k for k, g in enumerate(nodes)}
        mini_local = np.zeros((len(nodes), 4, 4), np.float32)
        mini_local[0] = np.eye(4, dtype=np.float32)
        mini_local[1:len(nodes) - 1024] = local_of(nodes[1:len(nodes) - 1024])
        mini_local[len(nodes) - 1024:] = scenes.random_objects(scenes.SEED_C3 + 1, int(leaves[0]), 1024)[3]
        mini_entries, mini_offsets = [], [0]
        for lvl in chain[::-1]:
            for g in lvl:
                mini_entries.append((remap[int(entries[int(g) - 1, 0])], remap[int(g)]))
            mini_offsets.append(len(mini_entries))
        mini_entries = np.asarray(mini_entries, np.uint32)
        mini_world = np.zeros_like(mini_local)
        mini_world[0] = np.eye(4, dtype=np.float32)
        nw = (len(nodes) + 31) // 32
        port.tree_compute(mini_local, mini_world, mini_entries, np.asarray(mini_offsets, np.uint32),
                          np.full(nw, 0xFFFFFFFF, np.uint32), np.zeros(nw, np.uint32))
        got_world = tree.world(int(leaves[0]), 1024)
        assert np.array_equal(got_world.view(np.uint32), mini_world[len(nodes) - 1024:].view(np.uint32)), "leaf world matrices at %d" % s
        lower4, extent4, _, _, _ = scenes.random_objects(scenes.SEED_C3, s, 1024)
        want = port.cull_bits(lower4, extent4, np.arange(1024, dtype=np.uint32), mini_world[len(nodes) - 1024:].reshape(-1), vp)
        assert np.array_equal(words[s // 32: s // 32 + 32], want), "visibility words at leaf %d" % s
    assert np.array_equal(r.changed(), _set_bits(~words, n))
    # the WHOLE frame against the oracle - Tree::compute over all 17.9 M nodes, then all 2^24 objects
    all_world = np.zeros_like(all_local)
    all_world[0] = np.eye(4, dtype=np.float32)
    nw_all = (n_nodes + 31) // 32
    port.tree_compute(all_local, all_world, entries, offsets, np.full(nw_all, 0xFFFFFFFF, np.uint32), np.zeros(nw_all, np.uint32))
    # C3 on several GPUs: the same frame as 4 shards - upper levels replicated, the leaf level sharded with
    # the objects, the leaf level still fused into the cull kernel, every shard's words stored into a "peer" full bitset
    from pipeline_b200 import sharding
    full = capi.Buffer(sharding.total_words(n) * 4)
    full.fill(0)
    for rank in range(4):
        e, o, nn, first_leaf_local, first, count, gmap = sharding.tree_shard(levels, 4, rank)
        st = capi.Tree(0)
        st.set_topology(e, o, nn)
        st.set_locals(0, np.ascontiguousarray(all_local[:first_leaf_local]))
        st.set_locals(first_leaf_local, np.ascontiguousarray(all_local[first_leaf + first:first_leaf + first + count]))
        st.mark_dirty(1, nn - 1)
        tix = capi.Buffer(count * 4)
        tix.upload(np.arange(first_leaf_local, nn, dtype=np.uint32))
        sctx = capi.Cull(0)
        sctx.set_objects(lo.ptr + first * 16, ex.ptr + first * 16, tix.ptr, capi.MEM_DEVICE, n=count)
        sr = sctx.result_create()
        sr.set_peer_bits([full.ptr], sharding.word_offset(first))
        sctx.run_with_tree(st, [sr], vp)
        assert sctx.get_option(capi.OPT_LAST_KERNEL) == capi.KERNEL_FUSED_LEAF
        assert np.array_equal(sr.bits(), words[first // 32:(first + count) // 32]), "shard %d" % rank
        sr.close(), sctx.close(), st.close(), tix.close()
    gathered = full.download(np.empty(sharding.total_words(n), np.uint32))
    assert np.array_equal(gathered, words), "gathered shard bitsets differ from the whole-tree cull"
    full.close()
    del all_local, gathered
    got_world = tree.world(1, n_nodes - 1)
    assert np.array_equal(got_world.view(np.uint32), all_world[1:].view(np.uint32)), "world matrices differ from the oracle"
    del got_world
    lo_h = lo.download(np.empty((n, 4), np.float32))
    ex_h = ex.download(np.empty((n, 4), np.float32))
    tidx_h = lo_h[:, 3].copy().view(np.uint32)
    assert np.array_equal(tidx_h, np.arange(first_leaf, n_nodes, dtype=np.uint32))
    lo_h[:, 3] = 1.0
    want_all = port.cull_bits(lo_h, ex_h, tidx_h, all_world.reshape(-1), vp, threads=8)
    bad = np.flatnonzero(words != want_all)
    assert bad.size == 0, "%d of %d visibility words differ from the oracle, first at word %d" % (bad.size, len(words), bad[0])
    del all_world, lo_h, ex_h, want_all
    # nothing dirty: no world matrix is recomputed, the published dirty set is empty, nothing changes
    ctx.run_with_tree(tree, [r], vp)
    assert r.changed_count() == 0
    assert _popcount(tree.dirty_world()) == 0
    assert np.array_equal(r.bits(), words)
    r.close(), ctx.close(), tree.close(), lo.close(), ex.close()
