#!/usr/bin/env python
"""bench.py - culled objects/s of the B200 culling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic input: a frustum cull of the
whole resident scene against a new camera, producing the visibility bitset and the ordered
changed-object list.  Under torchrun (N > 1) every rank owns a contiguous slice of the object
array on its own GPU (weak scaling: the slice size per GPU is fixed), there is no data-path
collective, and rank 0 prints ONE JSON line; `value` is the whole-job objects/s.

Workloads:
    c4-single   (default) 64 Mi random objects per GPU, one matrix each, single frustum - the
                configuration BASELINE.json's target is quoted on ("64M-object cull at 1 GPU").
    c2          1 Mi objects (fits in L2: an L2-flushing write runs between timed steps)
    c4          64 Mi objects x 6 cube-map frusta in one pass
    c3          16 Mi objects under a 4-level, 17.9 M-node transform hierarchy, every local
                matrix dirty every step: propagate (K1) + cull (K2) straight out of the tree
    c5          256 Mi objects strong-scaled over the N GPUs

Inputs are resident in HBM when the timed region starts (`value`); `e2e` is the same metric
through the C ABI with HOST buffers: per step the camera goes in and the bitset + changed list
come back into pinned host memory (host<->device copies inside the timed region).
`--impl reference` times the reference's own CPU culler (oracle/_ref, the unmodified
dp::culling::cpu::Manager) on the box's host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "culled objects/sec"
UNIT = "objects/s"

WORKLOADS = {
    #  name        objects/GPU  views  seed          scaling
    "c4-single": (1 << 26, 1, 0x5EED0004, "weak"),
    "c2":        (1 << 20, 1, 0x5EED0002, "weak"),
    "c4":        (1 << 26, 6, 0x5EED0004, "weak"),
    "c3":        (1 << 24, 1, 0x5EED0003, "weak"),
    "c5":        (1 << 28, 1, 0x5EED0005, "strong"),
}
C3_LEVELS = (4096, 65536, 1048576, 16777216)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load.  The sampler is
    started before the warm-up (nvidia-smi needs a few hundred ms to deliver its first line) and
    every line is stamped on arrival; the summary prefers the samples that fell inside the timed
    region and falls back to the whole loaded window (warm-up .. e2e) when the timed region was
    shorter than the sampling period - `window` says which."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []
        self.marks = {}

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 5.0:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line))

    def mark(self, name):
        self.marks[name] = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def parse(lo, hi):
            sm, mx, reasons, power = [], [], set(), []
            for t, line in self.lines:
                if not (lo <= t <= hi):
                    continue
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])), mx.append(float(f[2])), power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons, power

        m = self.marks
        window = "timed"
        sm, mx, reasons, power = parse(m.get("timed0", 0), m.get("timed1", 1e300))
        if len(sm) < 2:
            window = "load (warm-up..e2e; timed region shorter than the sampling period)"
            sm, mx, reasons, power = parse(m.get("load0", 0), m.get("load1", 1e300))
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU reference arm
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuCuller:
    """The reference CPU culler on a slice of the workload: oracle/_ref (kind "reference") when
    the compiled reference is present, else the C restatement (kind "port")."""

    def __init__(self, seed, first, count, threads):
        from oracle import loader
        from pipeline_b200 import scenes
        self.threads = threads
        self.count = count
        per = count // threads
        self.slices = []
        self.kind = "reference" if loader.Reference.available() else "port"
        ref = loader.Reference() if self.kind == "reference" else None
        port = loader.Port()
        self.port = port

        def setup(t):
            lower4, extent4, upper4, mats, tidx = scenes.random_objects(seed, first + t * per, per)
            tidx = np.arange(per, dtype=np.uint32)
            if ref is not None:
                s = ref.cull(0)
                s.add_objects(np.ascontiguousarray(lower4[:, :3]), np.ascontiguousarray(upper4[:, :3]), tidx)
                s.set_matrices(mats.reshape(-1))
                r = s.result_create()
                return (s, r, mats)
            return (lower4, extent4, tidx, mats.reshape(-1))

        ts, out = [], [None] * threads

        def work(t):
            out[t] = setup(t)
        for t in range(threads):
            th = threading.Thread(target=work, args=(t,))
            th.start()
            ts.append(th)
        for th in ts:
            th.join()
        self.slices = out
        self.total = per * threads

    def step(self, vp):
        """One cull of every slice, all threads concurrently; returns wall seconds."""
        def work(t):
            sl = self.slices[t]
            if self.kind == "reference":
                sl[0].cull(sl[1], vp)
            else:
                self.port.cull_bits(sl[0], sl[1], sl[2], sl[3], vp)
        ts = [threading.Thread(target=work, args=(t,)) for t in range(self.threads)]
        t0 = time.perf_counter()
        for th in ts:
            th.start()
        for th in ts:
            th.join()
        return time.perf_counter() - t0


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from pipeline_b200 import scenes
    n_per, views, seed, scaling = WORKLOADS[args.workload]
    cores = host_cores()
    threads = max(1, min(cores, 128))
    per_thread = 1 << 19
    culler = CpuCuller(seed, 0, per_thread * threads, threads)
    cams = [scenes.orbit_camera(f) for f in range(args.warmup + args.steps)]
    for f in range(args.warmup):
        culler.step(cams[f])
    secs = [culler.step(cams[args.warmup + f]) for f in range(args.steps)]
    ms = 1000.0 * sum(secs) / len(secs)
    value = culler.total / (ms / 1000.0)
    sample = ("%d objects (%d threads x %d, first objects of the %s scene, seed 0x%X), steady-state cull() per step, "
              "one %s per thread" % (culler.total, threads, per_thread, args.workload, seed,
                                     "dp::culling::cpu::Manager" if culler.kind == "reference" else "C port loop"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "objects_per_step": culler.total, "views": 1,
                   "note": "reference CPU culler on host cores; a step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": culler.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_single_core(workload, seed):
    """The reference as shipped is single-threaded (the omp pragma is commented out,
    dp/culling/cpu/src/ManagerImpl.cpp:510): 1 core, bounded sample, steady state."""
    from pipeline_b200 import scenes
    n = 1 << 22
    t0 = time.perf_counter()
    culler = CpuCuller(seed, 0, n, 1)
    setup_s = time.perf_counter() - t0
    cams = [scenes.orbit_camera(f) for f in range(12)]
    first = culler.step(cams[0])               # includes GroupCPU::updateOBBs
    culler.step(cams[1])
    secs = [culler.step(cams[2 + f]) for f in range(10)]
    med = float(np.median(secs))
    return {"value": n / med, "unit": UNIT, "cores": 1, "kind": culler.kind,
            "sample": "first %d objects of the %s scene (seed 0x%X), median of 10 steady-state cull() calls; "
                      "first cull incl. OBB build %.1f M objects/s; setup %.1f s" % (n, workload, seed, n / first / 1e6, setup_s)}


def cpu_tree_baseline(seed):
    """C3 only: dp::transform::Tree::compute on the host, every node dirty, on a 1/16-scale tree
    (256 / 4096 / 65536 / 1048576 nodes - building the full 17.9 M-node tree through addTransform takes minutes)."""
    from oracle import loader
    from pipeline_b200 import scenes
    levels = (256, 4096, 65536, 1048576)
    n_nodes = 1 + sum(levels)
    if loader.Reference.available():
        tree = loader.Reference().tree()
        parents = np.zeros(levels[0], np.uint32)
        first = 1
        for k, n in enumerate(levels):
            idx = tree.add_many(parents, scenes.random_objects(seed + 1, first, n)[3])
            first += n
            if k + 1 < len(levels):
                parents = np.repeat(idx, levels[k + 1] // n)
        secs = []
        for it in range(6):
            tree.update_locals(np.arange(1, n_nodes, dtype=np.uint32), scenes.random_objects(seed + 1, 1, n_nodes - 1)[3])
            t0 = time.perf_counter()
            tree.compute()
            secs.append(time.perf_counter() - t0)
        kind = "reference"
    else:
        port = loader.Port()
        entries, offsets, n_nodes = scenes.hierarchy_topology(levels)
        local = np.zeros((n_nodes, 4, 4), np.float32)
        local[0] = np.eye(4, dtype=np.float32)
        local[1:] = scenes.random_objects(seed + 1, 1, n_nodes - 1)[3]
        world = np.zeros_like(local)
        world[0] = local[0]
        nw = (n_nodes + 31) // 32
        secs = []
        for it in range(6):
            t0 = time.perf_counter()
            port.tree_compute(local, world, entries, offsets, np.full(nw, 0xFFFFFFFF, np.uint32), np.zeros(nw, np.uint32))
            secs.append(time.perf_counter() - t0)
        kind = "port"
    med = float(np.median(secs[1:]))
    return {"transform_nodes_per_s": (n_nodes - 1) / med, "transform_kind": kind,
            "transform_sample": "Tree::compute, %d nodes in 4 levels (1/16 of the C3 tree), all dirty, 1 core, median of 5" % (n_nodes - 1)}


# ------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    import torch
    from pipeline_b200 import capi, scenes

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    capi.device_select(device)
    torch.cuda.set_device(device)

    n_per, views, seed, scaling = WORKLOADS[args.workload]
    if scaling == "strong":
        n_per = n_per // world
    n_total = n_per * world
    first = rank * n_per
    W, K = args.warmup, args.steps
    if args.workload == "c2":
        K = max(K, 20)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        capi.device_sync()

    # ---------------- scene, resident in HBM
    stream = capi.Stream()
    tree = None
    c3_upper = 0
    if args.workload == "c3":
        from pipeline_b200 import sharding
        # SURVEY.md 8e "Transform propagation": levels 0..L-2 are replicated (every GPU propagates the 1.1 M upper nodes
        # itself, no communication), the leaf level is sharded with the objects.  Weak scaling: every GPU gets the full
        # 16 Mi leaves / objects, i.e. the tree a rank sees is the C3 tree with world x as many leaves in total.
        levels_total = C3_LEVELS[:-1] + (C3_LEVELS[-1] * world,)
        entries, offsets, n_nodes, c3_upper, first, cnt_leaves, _ = sharding.tree_shard(levels_total, world, rank)
        assert cnt_leaves == n_per
        tree = capi.Tree(device)
        if os.environ.get("DPCU_BENCH_TREE_WIDE_MIN"):          # developer A/B of the two K1 forms
            tree.set_option(capi.TREE_OPT_WIDE_MIN_NODES, int(os.environ["DPCU_BENCH_TREE_WIDE_MIN"]))
        tree.set_topology(entries, offsets, n_nodes)
        del entries
        first_leaf = n_nodes - n_per
        lo, ex = capi.Buffer(n_per * 16), capi.Buffer(n_per * 16)
        scratch = capi.Buffer(n_per * 64)
        # objects: random boxes; object i of this rank is bound to its leaf i (transformIndex = first_leaf + i)
        capi.scene_generate(seed, first, n_per, (first - first_leaf) & 0xFFFFFFFF, lo.ptr, ex.ptr, scratch.ptr)
        capi.device_sync()
        scratch.close()
        # local matrices: rigid transforms from the same generator, a function of the GLOBAL node index (bench input only)
        lptr, _ = tree.local_ptr()
        lo2, ex2 = capi.Buffer(n_nodes * 16), capi.Buffer(n_nodes * 16)
        capi.scene_generate(seed + 1, 0, c3_upper, 0, lo2.ptr, ex2.ptr, lptr)
        capi.scene_generate(seed + 1, c3_upper + first, n_per, 0, lo2.ptr, ex2.ptr, lptr + c3_upper * 64)
        capi.device_sync()
        lo2.close(), ex2.close()
        mats = None
    else:
        lo, ex, mats = capi.Buffer(n_per * 16), capi.Buffer(n_per * 16), capi.Buffer(n_per * 64)
        capi.scene_generate(seed, first, n_per, first, lo.ptr, ex.ptr, mats.ptr)
    capi.device_sync()

    ctx = capi.Cull(device)
    ctx.set_objects(lo.ptr, ex.ptr, None, capi.MEM_DEVICE, n=n_per)
    if tree is not None:
        wptr, cnt = tree.world_ptr()
        ctx.bind_matrices(wptr, cnt)
    else:
        ctx.bind_matrices(mats.ptr, n_per)
    ctx.set_option(capi.OPT_PROFILE, 1)
    results = [ctx.result_create() for _ in range(views)]

    # ---------------- multi-GPU: the bitset all-gather is the cull kernel's epilogue (NVLink peer stores).
    # The sharded cull itself has no exchange step; the all-gather exists for consumers
    # that need the full bitset on every GPU.  --gather also (default): the headline step is the cull
    # alone and the same step WITH the fused all-gather is measured afterwards and reported under
    # "also"; --gather fused: the all-gather is part of the headline step; --gather none: never.
    gather = "none"
    full_bits, peer_ptrs = [], []

    def enable_gather():
        nonlocal gather
        from pipeline_b200 import sharding
        try:
            words_total = sharding.total_words(n_total)
            for v in range(views):
                fb = capi.Buffer(words_total * 4)
                fb.fill(0)
                full_bits.append(fb)
                handles = sharding.exchange_ipc(dist, capi.ipc_get_handle(fb.ptr))
                ptrs = [fb.ptr if r == rank else capi.ipc_open(handles[r]) for r in range(world)]
                peer_ptrs.append(ptrs)
                results[v].set_peer_bits(ptrs, sharding.word_offset(first))
            gather = ("fused: every rank's cull kernel (cullLinesKernel) stores its finished 128-byte lines into all %d full "
                      "bitsets (cudaIpc peer pointers, NVLink stores)" % world)
            return True
        except Exception as e:                      # noqa: BLE001 - report, do not hide
            gather = "none (peer mapping failed: %s)" % e
            for v in range(views):
                results[v].set_peer_bits([], 0)
            return False

    def verify_gather():
        from pipeline_b200 import sharding
        barrier()
        local = torch.from_numpy(results[0].bits().view(np.int32)).cuda()
        parts = sharding.allgather_words(dist, local, n_words)
        want = torch.cat(parts).cpu().numpy().view(np.uint32)
        got = np.zeros(sharding.total_words(n_total), np.uint32)
        full_bits[0].download(got)
        return bool(np.array_equal(got, want[:len(got)]))

    if world > 1 and args.gather == "fused":
        enable_gather()

    if views > 1:
        # six static faces plus a slowly translating eye so that the changed lists are not empty
        cams = []
        for f in range(W + 2 * K + 4):
            eye = (3.0 * f, 0.0, 0.0)
            cams.append(np.ascontiguousarray(scenes.cube_map_cameras(eye)[:views], np.float32))
    else:
        cams = [np.ascontiguousarray(scenes.orbit_camera(f), np.float32).reshape(1, 16) for f in range(W + 2 * K + 4)]

    flush = sweep = None
    if n_per * 96 < 200e6:                     # working set fits in the 126 MB L2: flush between timed steps
        flush = capi.Buffer(256 << 20)
        sweep = capi.Buffer(256 << 20)
        sweep.fill(1)

    # per-launch events inside the timed region only where the cull kernel is the step's only launch (large groups); for the
    # flushed small workloads and the tree-fed frame they would sit between the cull kernel and the dependent launches around
    # it (tree levels before, compaction kernel behind) - see kernel_time_source below
    profile_in_timed = flush is None and tree is None
    ctx.set_option(capi.OPT_PROFILE, 1 if profile_in_timed else 0)

    def step(f, s=stream):
        if tree is not None:
            tree.mark_dirty(1, tree.n_nodes - 1)   # every local matrix "rewritten" this frame
            ctx.run_with_tree(tree, results, cams[f], s)   # levels 0..2 propagate, the leaf level runs inside the cull kernel
        else:
            ctx.run(results, cams[f], s)

    sampler = ClockSampler(device)
    sampler.start()
    sampler.mark("load0")
    for f in range(W):
        step(f)
    stream.sync()
    ctx.kernel_time()
    launches0 = ctx.launches() + (tree.launches() if tree else 0)

    # ---------------- timed region: K steps, device-resident inputs
    e0, e1 = capi.Event(), capi.Event()
    barrier()
    sampler.mark("timed0")
    if flush is None:
        e0.record(stream)
        for f in range(K):
            step(W + f)
        e1.record(stream)
        stream.sync()
        ms_total = e0.elapsed_ms(e1)
    else:
        # Small workload: every step is bracketed by its own event pair, and in front of it the L2 is flushed - 256 MiB
        # written (> L2), then another 256 MiB read, so that the L2 is cold AND clean (the write-back of the flush's own
        # dirty lines must not land inside the timed step).  Everything is queued on the stream without a host
        # synchronisation in between: the flush keeps the GPU busy while the next step is enqueued, so the event pair
        # measures device time of the step, not the host's launch overhead.
        evs = [(capi.Event(), capi.Event()) for _ in range(K)]
        for f in range(K):
            flush.fill(f & 0xFF, stream)
            capi.read_sweep(sweep.ptr, 256 << 20, stream)
            evs[f][0].record(stream)
            step(W + f)
            evs[f][1].record(stream)
        stream.sync()
        ms_total = float(sum(x.elapsed_ms(y) for x, y in evs))
    barrier()
    sampler.mark("timed1")
    launches1 = ctx.launches() + (tree.launches() if tree else 0)
    headline_kernel = ctx.get_option(capi.OPT_LAST_KERNEL)       # which exact form AUTO picked for the timed steps
    kernel_time_source = "CUDA events around every cull-kernel launch of the timed region (DPCU_CULL_OPT_PROFILE)"
    if not profile_in_timed:
        # separate pass, same protocol: the events would sit between the cull kernel and its dependent compaction launch
        ctx.set_option(capi.OPT_PROFILE, 1)
        ctx.kernel_time()
        for f in range(K):
            if flush is not None:
                flush.fill(f & 0xFF, stream)
                capi.read_sweep(sweep.ptr, 256 << 20, stream)
            step(W + f)
        stream.sync()
        kernel_time_source = ("CUDA events around every cull-kernel launch in a SEPARATE pass of the same %d steps (same protocol) right "
                              "after the timed region: events between the cull kernel and the programmatic dependent launches around it "
                              "serialise them (measured: +7 us per step at 1 Mi objects), so the timed steps run without them" % K)
    k_each = ctx.kernel_times()                                  # per-launch device time of the cull kernel (CUDA events)
    ctx.set_option(capi.OPT_PROFILE, 1 if profile_in_timed else 0)
    cold_read = None
    if flush is not None and tree is None:
        # what a launch of this size can reach at all: a plain streaming read of the same number of bytes (nothing computed,
        # nothing written), same flush protocol, same event bracket - launch latency, ramp-up and the first DRAM round trips
        # are in it just as they are in the cull kernel's time; the copy peak of MEASURED_PEAKS.json is a large-transfer figure
        rb = capi.Buffer(int(n_per * 96))
        rb.fill(3)
        evr = [(capi.Event(), capi.Event()) for _ in range(K)]
        for f in range(K):
            flush.fill(f & 0xFF, stream)
            capi.read_sweep(sweep.ptr, 256 << 20, stream)
            evr[f][0].record(stream)
            capi.read_sweep(rb.ptr, int(n_per * 96), stream)
            evr[f][1].record(stream)
        stream.sync()
        tr = np.array([x.elapsed_ms(y) for x, y in evr])
        cold_read = {"bytes": int(n_per * 96), "ms_median": float(np.median(tr)), "ms_min": float(tr.min()),
                     "GBps": n_per * 96 / float(np.median(tr)) / 1e6,
                     "what": "readSweepKernel (grid-stride 16-byte loads, result discarded) over the same number of bytes the cull reads, "
                             "cold clean L2, one event pair per launch: the floor for ANY single launch of this size on this GPU"}
        rb.close()
    k_ms, k_n = float(k_each.sum()), len(k_each)
    changed = [r.changed_count() for r in results]

    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / K
    value = n_total / (ms_per_step / 1000.0)

    # ---------------- e2e: same steps through the C ABI with host buffers.  The result is mirrored in pinned
    # host memory (dpcuCullResultSetHostMirror): the cull kernel's epilogue stores whole bitset lines and the
    # compaction kernel the changed list over PCIe while they run; the step ends with dpcuCullResultSynchronize.
    n_words = (n_per + 31) // 32
    pinned_bits = [capi.HostBuffer(n_words * 4) for _ in range(views)]
    pinned_chg = [capi.HostBuffer(max(n_per, 1) * 4) for _ in range(views)]
    pinned_cnt = [capi.HostBuffer(4) for _ in range(views)]
    hb = [p.array(np.uint32) for p in pinned_bits]
    hc = [p.array(np.uint32) for p in pinned_chg]
    hn = [p.array(np.uint32) for p in pinned_cnt]
    pinned_vp = capi.HostBuffer(views * 64)
    hvp = pinned_vp.array(np.float32)
    import ctypes as C
    L = capi.lib()
    u32p = C.POINTER(C.c_uint32)

    def e2e_step(f, mirrored=True):
        hvp[:] = cams[f].reshape(-1)           # the step's input lives in pinned host memory
        if tree is not None:
            tree.mark_dirty(1, tree.n_nodes - 1)
            ctx.run_with_tree(tree, results, hvp, stream)
        else:
            ctx.run(results, hvp, stream)
        d2h = 0
        for v in range(views):
            if mirrored:
                results[v].synchronize()
                d2h += n_words * 4 + 4 + int(hn[v][0]) * 4
            else:
                capi.check(L.dpcuCullResultGetBits(results[v].h, hb[v].ctypes.data_as(u32p), n_words))
                cnt = C.c_size_t()
                capi.check(L.dpcuCullResultGetChanged(results[v].h, hc[v].ctypes.data_as(u32p), n_per, C.byref(cnt)))
                d2h += n_words * 4 + 4 + cnt.value * 4
        return d2h

    def e2e_run(mirrored):
        for f in range(2):
            e2e_step(W + K + f, mirrored)
        barrier()
        nbytes = 0
        t0 = time.perf_counter()
        e0.record(stream)
        for f in range(K):
            nbytes += e2e_step(W + K + 2 + f, mirrored)
        e1.record(stream)
        stream.sync()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        ms = max(e0.elapsed_ms(e1), wall_ms)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, nbytes

    copy_ms, _ = e2e_run(False)                # the copy-after-the-cull getters, for comparison (reported under "also")
    for v in range(views):
        results[v].set_host_mirror(hb[v], hc[v], hn[v])
    e2e_ms, d2h_bytes = e2e_run(True)
    # the mirror really is the result: compare with the device-side getters once, untimed
    e2e_checked = all(np.array_equal(hb[v][:n_words], results[v].bits()) and int(hn[v][0]) == results[v].changed_count()
                      and np.array_equal(hc[v][:int(hn[v][0])], results[v].changed()) for v in range(views))
    for v in range(views):
        results[v].set_host_mirror(None, None, None)
    e2e_value = n_total / (e2e_ms / K / 1000.0)
    ctx.kernel_time()
    sampler.mark("load1")
    clocks = sampler.stop()

    # ---------------- multi-GPU: check the fused all-gather against the library collective (untimed)
    gather_ok = None
    if full_bits:
        gather_ok = verify_gather()

    # ---------------- roofline of the dominant kernel (K2, the cull kernel)
    peak, peak_src = measured_peak()
    alg_bytes = n_per * (96.0 + 0.25 * views)            # per launch, per GPU
    kernel_name = "%s<%d>" % (capi.KERNEL_NAMES.get(headline_kernel, "cullDirectKernel"), views)
    if tree is not None:
        # fused leaf level: local 64 + entry 8 + world write 64 + AABB 32 + bits, parents (1/16 of the leaves) 64 each
        alg_bytes = n_per * (64.0 + 8.0 + 64.0 + 32.0 + 0.25 * views) + C3_LEVELS[-2] * 64.0
    k_avg_ms = k_ms / max(k_n, 1)
    achieved = alg_bytes / (k_avg_ms / 1000.0) / 1e9
    traffic = None                                       # dram bytes of the same kernel / workload from the committed ncu capture
    traffic_source = None
    for summary in ("r02_cull_kernel_summary.json", "r01_cull_kernel_summary.json"):
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", summary)))
        except Exception:
            continue
        for p in prof.get("captures", []):
            if (traffic is None and p.get("workload") == args.workload and p.get("objects") == n_per and p.get("views") == views
                    and kernel_name.split("<")[0] in p.get("kernel", "")):
                traffic = p["dram_bytes_read"] + p["dram_bytes_write"]
                traffic_source = "profiles/%s (%s): ncu --set full capture of this kernel on this workload, dram__bytes_read.sum + dram__bytes_write.sum per launch" % (summary, p.get("report", "?"))
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k_avg_ms, "launches_timed": int(k_n),
                "launch_ms_p10_p50_p90": [float(x) for x in np.percentile(k_each, [10, 50, 90])] if len(k_each) else None,
                "step_share": k_ms / ms_total if ms_total else None, "kernel_time_source": kernel_time_source}
    if cold_read is not None:
        roofline["cold_read_floor"] = cold_read
        roofline["frac_of_cold_read_floor"] = cold_read["ms_median"] / k_avg_ms if k_avg_ms else None
    if tree is not None:
        upper = sum(C3_LEVELS[:-1])
        alg_upper = upper * 136.0 + (1 + sum(C3_LEVELS[:-2])) * 64.0     # levels 0..2: K1 launches
        roofline["step_algorithmic_bytes"] = alg_upper + alg_bytes       # 3.05 GB: the fused figure of SURVEY.md 8d
        roofline["step_achieved"] = (alg_upper + alg_bytes) / (ms_per_step / 1000.0) / 1e9
        roofline["step_frac"] = roofline["step_achieved"] / peak

    also = {"e2e_via_copy_getters": {"ms_per_step": copy_ms / K, "objects_per_s": n_total / (copy_ms / K / 1000.0),
                                     "what": "dpcuCullRun, then dpcuCullResultGetBits / GetChanged copies (no host mirror)"}}
    # ---------------- the same step with the bitset all-gather fused into the cull kernel (N > 1)
    if world > 1 and args.gather == "also":
        if enable_gather():
            for f in range(3):
                step(f)
            stream.sync()
            ctx.kernel_time()
            barrier()
            e0.record(stream)
            for f in range(K):
                step(W + f)
            e1.record(stream)
            stream.sync()
            ms_g = e0.elapsed_ms(e1)
            t = torch.tensor([ms_g], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_g = float(t.item()) / K
            kg_ms, kg_n = ctx.kernel_time()
            ok = verify_gather()
            also["with_bitset_allgather"] = {
                "ms_per_step": ms_g, "objects_per_s": n_total / (ms_g / 1000.0),
                "kernel": "%s<%d>%s" % (capi.KERNEL_NAMES.get(ctx.get_option(capi.OPT_LAST_KERNEL), "?"), views,
                                        " + peerGatherKernel" if tree is not None else ""),
                "avg_launch_ms": kg_ms / max(kg_n, 1), "how": gather, "verified_against_nccl_all_gather": ok,
                "nvlink_bytes_out_per_gpu_per_step": (world - 1) * n_words * 4 * views}
            for v in range(views):
                results[v].set_peer_bits([], 0)
        gather = "not part of the headline step (no data-path collective); measured separately under also.with_bitset_allgather"
        gather_ok = None

    # ---------------- K4: the group bounding box over the same resident scene (same 96 B/object stream)
    if tree is None and not args.no_also:
        ctx.bounding_box()
        capi.device_sync()
        t_bb = []
        for _ in range(5):
            t0 = time.perf_counter()
            box = ctx.bounding_box()
            t_bb.append(time.perf_counter() - t0)
        bb_ms = 1000.0 * float(np.median(t_bb))
        also["bounding_box"] = {
            "ms": bb_ms, "objects_per_s": n_per / (bb_ms / 1000.0), "achieved_GBps": n_per * 96.0 / (bb_ms / 1000.0) / 1e9,
            # (a small group's synchronous call is launch + copy latency, not bandwidth: no fraction is claimed for it)
            "frac_of_hbm_peak": (n_per * 96.0 / (bb_ms / 1000.0) / 1e9 / peak) if flush is None else None, "box": [float(x) for x in box],
            "what": "dpcuCullGetBoundingBox (ManagerBitSet::calculateBoundingBox): one kernel over the object stream + the 24-byte "
                    "read-back, wall clock of the synchronous call on this rank's slice"}

    # ---------------- FMA fast mode: reporting only (north_star: its boundary-object disagreements are reported separately)
    if tree is None and rank == 0 and not args.no_also:
        rep = ctx.fma_report(cams[W + K][0])
        also["fma_mode"] = {"objects": rep["objects"], "disagreements": rep["disagreements"], "first_indices": rep["first_indices"],
                            "what": "the same cull with the direct kernel compiled -fmad=true (DPCU_CULL_OPT_FMA = 1) against the exact "
                                    "product path, same camera: objects whose visibility bit differs; never used for results"}

    # ---------------- BASELINE config 5 next to the headline: 256 Mi objects STRONG-scaled over the N GPUs, with the bitset
    # all-gather fused into the timed step when N > 1 (the driver's 1/2/4/8 sweep then shows the c5 strong-scaling curve)
    if args.workload == "c4-single" and not args.no_also and not args.no_c5:
        c5_total = int(os.environ.get("DPCU_BENCH_C5_TOTAL", 1 << 28))      # developer override (e.g. 2 GPUs at 8-GPU shard size)
        n5 = c5_total // world
        first5 = rank * n5
        lo5, ex5, mt5 = capi.Buffer(n5 * 16), capi.Buffer(n5 * 16), capi.Buffer(n5 * 64)
        capi.scene_generate(WORKLOADS["c5"][2], first5, n5, first5, lo5.ptr, ex5.ptr, mt5.ptr)
        capi.device_sync()
        ctx5 = capi.Cull(device)
        ctx5.set_objects(lo5.ptr, ex5.ptr, None, capi.MEM_DEVICE, n=n5)
        ctx5.bind_matrices(mt5.ptr, n5)
        ctx5.set_option(capi.OPT_PROFILE, 1)
        if os.environ.get("DPCU_BENCH_LINE_WORDS"):
            ctx5.set_option(capi.OPT_LINE_WORDS, int(os.environ["DPCU_BENCH_LINE_WORDS"]))
        r5 = ctx5.result_create()
        g5 = "none (1 GPU)"
        fb5 = None
        ok5 = None
        if world > 1:
            from pipeline_b200 import sharding
            fb5 = capi.Buffer(sharding.total_words(c5_total) * 4)
            fb5.fill(0)
            handles = sharding.exchange_ipc(dist, capi.ipc_get_handle(fb5.ptr))
            ptrs5 = [fb5.ptr if r == rank else capi.ipc_open(handles[r]) for r in range(world)]
            r5.set_peer_bits(ptrs5, sharding.word_offset(first5))
            g5 = "fused into the cull kernel's epilogue: NVLink peer stores of finished 128-byte lines into all %d full bitsets" % world
        for f in range(W):
            ctx5.run([r5], cams[f], stream)
        stream.sync()
        ctx5.kernel_time()
        barrier()
        e0.record(stream)
        for f in range(10):
            ctx5.run([r5], cams[W + f], stream)
        e1.record(stream)
        stream.sync()
        ms5 = e0.elapsed_ms(e1)
        if dist is not None:
            t = torch.tensor([ms5], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms5 = float(t.item())
        ms5 /= 10
        k5_ms, k5_n = ctx5.kernel_time()
        if world > 1:
            barrier()
            local5 = torch.from_numpy(r5.bits().view(np.int32)).cuda()
            parts = sharding.allgather_words(dist, local5, (n5 + 31) // 32)
            want5 = torch.cat(parts).cpu().numpy().view(np.uint32)
            got5 = np.zeros(sharding.total_words(c5_total), np.uint32)
            fb5.download(got5)
            ok5 = bool(np.array_equal(got5, want5[:len(got5)]))
            del local5, parts, want5, got5
        alg5 = n5 * 96.25
        also["c5_strong"] = {
            "objects_total": c5_total, "objects_per_gpu": n5, "scaling": "strong", "ms_per_step": ms5,
            "objects_per_s": c5_total / (ms5 / 1000.0),
            "kernel": "%s<1>" % capi.KERNEL_NAMES.get(ctx5.get_option(capi.OPT_LAST_KERNEL), "?"), "avg_launch_ms": k5_ms / max(k5_n, 1),
            "frac_of_hbm_peak": alg5 / (k5_ms / max(k5_n, 1) / 1000.0) / 1e9 / peak,
            "bitset_allgather": g5, "bitset_allgather_verified_against_nccl_all_gather": ok5,
            "nvlink_bytes_out_per_gpu_per_step": (world - 1) * ((n5 + 31) // 32) * 4,
            "limiter": "1 GPU: HBM; N > 1: the cull kernel's peer-store epilogue (every finished line goes out N - 1 times over NVLink)"}
        r5.close()
        ctx5.close()
        for b in (lo5, ex5, mt5):
            b.close()

    # ---------------- the same resident scene against six cube-map frusta in one pass (BASELINE config 4)
    if args.workload == "c4-single" and not args.no_also:
        res6 = [ctx.result_create() for _ in range(6)]
        cams6 = [np.ascontiguousarray(scenes.cube_map_cameras((3.0 * f, 0.0, 0.0)), np.float32) for f in range(W + 10)]
        for f in range(W):
            ctx.run(res6, cams6[f], stream)
        stream.sync()
        ctx.kernel_time()
        barrier()
        e0.record(stream)
        for f in range(10):
            ctx.run(res6, cams6[W + f], stream)
        e1.record(stream)
        stream.sync()
        ms6 = e0.elapsed_ms(e1)
        if dist is not None:
            t = torch.tensor([ms6], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms6 = float(t.item())
        ms6 /= 10
        k6_ms, k6_n = ctx.kernel_time()
        alg6 = n_per * (96.0 + 0.25 * 6)
        also["c4_six_views"] = {
            "ms_per_step": ms6, "objects_per_s": n_total / (ms6 / 1000.0), "object_views_per_s": 6 * n_total / (ms6 / 1000.0),
            "kernel": "%s<6>" % capi.KERNEL_NAMES.get(ctx.get_option(capi.OPT_LAST_KERNEL), "?"), "avg_launch_ms": k6_ms / max(k6_n, 1),
            "achieved_GBps": alg6 / (k6_ms / max(k6_n, 1) / 1000.0) / 1e9, "frac_of_hbm_peak": alg6 / (k6_ms / max(k6_n, 1) / 1000.0) / 1e9 / peak,
            "bound": "hbm (0.82: one exposed DRAM round trip per step at 8 warps per scheduler; provable pairs decided by the "
                     "centre / radius filter two views per instruction, the rest queued for the reference arithmetic; DESIGN.md section 3)",
            "changed_per_step_last": [r.changed_count() for r in res6]}
        for r in res6:
            r.close()

    # ---------------- the other single-GPU BASELINE configurations next to the headline (N = 1 only): the same bench, one
    # child process per workload (its own context, its own flush protocol), summarised here so that the record of the
    # default command carries every named configuration; the full lines are what `bench.py --workload c2|c3` prints
    if args.workload == "c4-single" and world == 1 and not args.no_also and not args.no_children:
        import subprocess
        for wl in ("c2", "c3"):
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", wl, "--no-also", "--no-cpu-baseline",
                                      "--steps", "20", "--warmup", "5"], capture_output=True, text=True, timeout=300)
                child = json.loads(out.stdout.strip().splitlines()[-1])
                cr = child["roofline"]
                also[wl] = {"config": {k: child["config"][k] for k in ("workload", "objects_total", "views", "matrices", "l2")},
                            "ms_per_step": child["ms_per_step"], "objects_per_s": child["value"],
                            "e2e_ms_per_step": child["e2e"]["ms_per_step"], "e2e_objects_per_s": child["e2e"]["value"],
                            "gpu_launches": child["gpu_launches"], "steps": child["steps"],
                            "roofline": {k: cr.get(k) for k in ("kernel", "achieved", "peak", "frac", "avg_launch_ms", "step_share", "traffic",
                                                                "algorithmic_bytes_per_launch", "step_frac", "kernel_time_source",
                                                                "cold_read_floor", "frac_of_cold_read_floor") if k in cr}}
            except Exception as e:                   # noqa: BLE001 - the headline line must not depend on the extras
                also[wl] = {"failed": "%s: %s" % (type(e).__name__, e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "objects_per_gpu": n_per, "objects_total": n_total, "views": views,
                   "matrices": "one per object (transformIndex = i)" if tree is None else "4-level tree, %d nodes" % tree.n_nodes,
                   "camera": "new view-projection every step", "changed_per_step_last": changed,
                   "l2": "inputs (%.1f GB per GPU) larger than L2" % (n_per * 96 / 1e9) if flush is None
                         else "before every timed step: 256 MiB flush write, then a 256 MiB read sweep (cold, clean L2); one event pair per step",
                   "parallelism": "object slices, one per GPU, no data-path collective" if world > 1 else "single GPU",
                   "bitset_allgather": gather, "bitset_allgather_verified": gather_ok,
                   "exact_mode": "-fmad=false, bit-exact vs dp::culling::cpu"},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": views * 64, "d2h_bytes_per_step": d2h_bytes // K,
                "ms_per_step": e2e_ms / K,
                "mirror_equals_device_result": e2e_checked,
                "what": "camera from pinned host memory in, visibility bitset + changed list into pinned host memory out, "
                        "through dpcuCullRun / dpcuCullResultSynchronize with a host mirror (dpcuCullResultSetHostMirror): "
                        "the cull kernel stores bitset lines and changed-list runs over PCIe while it runs, no copy after the cull"},
        "gpu_launches": int(launches1 - launches0),
        "clocks": clocks,
    }
    if also:
        line["also"] = also
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single_core(args.workload, seed)
        if tree is not None:
            line["cpu_baseline"].update(cpu_tree_baseline(seed))
    if rank == 0:
        print(json.dumps(line), flush=True)
    for r in results:
        r.close()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4-single", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the extra measurements over the same scene (six views, c5 strong, K4, FMA report)")
    ap.add_argument("--no-children", action="store_true", help="skip also.c2 / also.c3 (child runs of the other single-GPU workloads)")
    ap.add_argument("--no-c5", action="store_true", help="skip also.c5_strong (256 Mi objects over the N GPUs)")
    ap.add_argument("--gather", default="also", choices=["also", "fused", "none"],
                    help="N>1: bitset all-gather through peer stores in the cull kernel's epilogue: measured next to the "
                         "headline step (also, default), inside it (fused), or not at all (none)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
