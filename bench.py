#!/usr/bin/env python
"""Benchmark of the per-bin reconstruction hot path (octree build -> MLS -> marching + weld).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): a 10M-splat synthetic bumpy torus with
per-splat normals and radii, --fit-grid 0.002, sphere fit, CLI defaults (smooth 4, levels 6,
subsampling 3, bins of <=255 cells).  One "step" = one pass of the hot path over all bins.
Under torchrun (N ranks) the ranks mesh ONE torus of N x 10M splats at fit-grid 0.002 / sqrt(N) (the same
splats per cell, N times the bins): its bins are dealt to the ranks in contiguous angular runs of
equal load (mlsgpu_b200.host.ShardedTorus), every rank builds and processes only its share — weak
scaling, no collective on the data path; rank 0 prints one JSON line.  --replicas restores the
round-1 behaviour (every rank its own 10M torus).

value   = splats/s with the bins' splats resident in HBM at the start of the timed region
e2e     = splats/s through the C ABI from pinned HOST buffers (H2D of every bin and D2H of every
          mesh inside the timed region)
roofline= processCorners against the FP32 pipe: (7 T + 16 H) pipe instructions over
          the kernel's CUDA-event time, against a measured dependent-free FFMA stream;
          rooflines_hbm = octree and marching against the measured HBM copy peak (model bytes of
          SURVEY 8d over the stages' kernel time)
parity_checked = after the timed regions two bins of the timed workload (the largest and a random
          one) are meshed once more and compared with the CPU oracle: vertices, keys, triangles
cpu_baseline = the CPU oracle (port of the reference kernels) on a bounded sample of the bins
"""
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

FIT_GRID = 0.002
SMOOTH = 4.0
PROFILE_JSON = "r2_kernels_ncu.json"
MLS_STAT = "kernel.mls.processCorners.time"
OCTREE_STATS = ["kernel.octree.writeEntries.time", "kernel.octree.sort.time", "kernel.octree.fill.time",
                "kernel.octree.countCommands.time", "kernel.octree.scan.time", "kernel.octree.writeSplatIds.time",
                "kernel.octree.writeStart.time"]
MARCHING_STATS = ["kernel.marching.zero.time", "kernel.marching.genOccupied.time", "kernel.marching.scanElements.time",
                  "kernel.marching.countUniqueVertices.time", "kernel.marching.compactVertices.time",
                  "kernel.marching.generateElements.time", "kernel.marching.readback.time"]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_workload(num_splats, seed, rank=0, world=1):
    """world == 1: C2 itself.  world > 1: this rank's share of ONE torus of world x num_splats splats."""
    from mlsgpu_b200 import host
    t0 = time.time()
    if world > 1:
        st = host.ShardedTorus(num_splats * world, world, FIT_GRID / np.sqrt(world), seed=seed, smooth=SMOOTH)
        g, bins, lo, spacing = st.rank_share(rank)
        total_splats = st.n
    else:
        s = host.torus_splats(num_splats, seed=seed)
        g, bins, lo, spacing = host.prepare(s, FIT_GRID, smooth=SMOOTH)
        total_splats = num_splats
    # concatenate the bins' splats: this is what the loader hands to the copy stage
    firsts, total = [], 0
    for b in bins:
        firsts.append(total)
        total += len(b["ids"])
    packed = np.empty(total, g.dtype)
    for b, f in zip(bins, firsts):
        packed[f:f + len(b["ids"])] = g[b["ids"]]
        b["first"] = f
        b["n"] = len(b["ids"])
    log("workload (rank %d of %d): %d splats in all, %d bins here, %d splat instances, prepared in %.1fs"
        % (rank, world, total_splats, len(bins), total, time.time() - t0))
    return dict(splats=packed, bins=bins, lo=lo, spacing=spacing, num_splats=total_splats, world=world)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc = None
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = self.proc.communicate()[0]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class Runner:
    """`threads` workers on one GPU (the reference's --device-threads), bins handed out
    round-robin; each worker thread uploads its bin and runs mlsb_worker_process_bin."""

    def __init__(self, wl, device, threads, order="largest-first"):
        import mlsgpu_b200 as m
        from mlsgpu_b200.capi import DeviceBuffer, PinnedBuffer, MESH_SINK
        self.m = m
        self.wl = wl
        self.bins = wl["bins"]
        total = len(wl["splats"])
        self.max_bin = max(b["n"] for b in self.bins)
        # the workers take bins from one queue; handing out the big ones first keeps the end of a
        # step from waiting on a single large bin (the reference's copy stage is free to pick, too)
        self.order = sorted(range(len(self.bins)), key=lambda i: -self.bins[i]["n"]) if order == "largest-first" \
            else list(range(len(self.bins)))
        # the native worker group (CopyGroup + DeviceWorkerGroup): `threads` worker threads, bins are
        # copied straight from the pointer given to submit (pinned host memory or HBM)
        self.group = m.WorkerGroup(num_workers=threads, stage=False, sink=None, device=device,
                                   max_splats=self.max_bin, max_cells=255, mesh_bytes=256 << 20)
        self.workers = self.group.workers
        # a lone worker for the per-kernel roofline pass (no other kernels on the GPU)
        self.single = m.Worker(device=device, max_splats=self.max_bin, max_cells=255, mesh_bytes=256 << 20)
        bias = wl["lo"].astype(np.float32) * wl["spacing"]
        for w in self.workers + [self.single]:
            w.set_scale_bias(float(wl["spacing"]), bias)
        self.pinned = PinnedBuffer(m.SPLAT_DTYPE, total)
        self.pinned.array[:] = wl["splats"]
        self.resident = DeviceBuffer.from_array(wl["splats"])
        self.mesh_bytes = 0
        self.mesh_vertices = 0
        self.mesh_triangles = 0
        self.lock = threading.Lock()

        def sink(sizes_p, blob, user):
            s = sizes_p.contents
            with self.lock:
                self.mesh_vertices += int(s.numVertices)
                self.mesh_triangles += int(s.numTriangles)
                self.mesh_bytes += 12 * int(s.numVertices) + 12 * int(s.numTriangles) + 8 * int(s.numVertices - s.numInternalVertices)
        self.sink = MESH_SINK(sink)

    def step(self, from_host, single=False):
        """One pass over all bins.  single=True: the lone worker (kernels do not overlap, which is
        what a per-kernel roofline measurement needs); otherwise the worker group."""
        from mlsgpu_b200.capi import lib, _check
        base = self.pinned.ptr.value if from_host else self.resident.ptr.value
        if single:
            L, w = lib(), self.single
            for b in self.bins:
                _check(L.mlsb_splats_upload_async(w.h, base + 32 * b["first"], 0, b["n"]))
                _check(L.mlsb_worker_process_bin(w.h, 0, b["n"], b["size"].ctypes.data, b["offset"].ctypes.data,
                                                 self.sink, None))
            return
        before = self.group.counters()[2]
        for i in self.order:
            b = self.bins[i]
            self.group.submit((base + 32 * b["first"], b["n"]), b["size"], b["offset"], tag=i)
        self.group.wait()
        self.mesh_bytes += self.group.counters()[2] - before

    def sync(self):
        self.group.wait()
        for w in self.workers + [self.single]:
            w.synchronize()

    def all_workers(self):
        return self.workers + [self.single]

    def blob_bytes_since_reset(self):
        """HostKeyMesh bytes of every block shipped since the workers' statistics were reset (for the wire / blob ratio)."""
        v = sum(w.marching_work()[4] for w in self.all_workers())       # welded vertices
        i = sum(w.marching_work()[3] for w in self.all_workers())       # indices
        return 12 * v + 4 * i    # external keys (8 bytes each, < 1 % of a block) are left out on both sides of the ratio

    def stat_sum(self, names):
        tot, cnt = 0.0, 0
        for w in self.all_workers():
            for n in names:
                s, c = w.statistic(n)
                tot += s
                cnt += c
        return tot, cnt

    def close(self):
        self.group.close()
        self.single.close()
        self.resident.free()
        self.pinned.free()


def oracle_bin(orc, wl, b):
    """The reference path for one bin on the CPU oracle: build -> processCorners -> marching."""
    s = wl["splats"][b["first"]:b["first"] + b["n"]].copy()
    tree = orc.octree_build(s, 0, len(s), b["offset"], 3, 6)
    size = [int(v) for v in b["size"]]
    pitch, zs, depth = (-(-v // 8) * 8 for v in size)
    img = np.zeros((zs * depth, pitch), np.float32)
    orc.mls_process_corners(img, s, tree["commands"], tree["start"], 9, b["offset"], size[0], size[1], zs, 0, 0,
                            size[2] - 1, orc.boundary_factor(1.0), 0)
    mesh = orc.marching_generate(img, size[0], size[1], zs, 0, 0, size[2] - 1, [int(v) for v in b["offset"]],
                                 vcap=4 << 20, icap=24 << 20)
    return mesh


def parity_check(runner, wl, seed):
    """Untimed: the largest bin and a random one of the timed workload, meshed by the lone worker and
    compared with the CPU oracle (vertices after scale/bias, external keys, internal split, triangles)."""
    from oracle.oracle import Oracle
    orc = Oracle()
    order = sorted(range(len(wl["bins"])), key=lambda i: -wl["bins"][i]["n"])
    rng = np.random.default_rng(seed)
    picks = [order[0], int(rng.choice(order[1:max(2, len(order) * 3 // 4)]))]
    w = runner.single
    scale = np.float64(np.float32(wl["spacing"]))
    bias = (wl["lo"].astype(np.float32) * wl["spacing"]).astype(np.float64)
    identical, nv = True, 0
    for i in picks:
        b = wl["bins"][i]
        s = np.ascontiguousarray(wl["splats"][b["first"]:b["first"] + b["n"]])
        w.upload_splats(s)
        blocks = w.process_bin(0, b["n"], b["size"], b["offset"])
        ref = oracle_bin(orc, wl, b)
        ok = len(blocks) == (1 if len(ref["vertices"]) else 0)
        if ok and blocks:
            g = blocks[0]
            want = (ref["vertices"].astype(np.float64) * scale + bias).astype(np.float32)   # fma: one rounding
            ni = ref["num_internal"]
            ok = (g["vertices"].shape == want.shape and np.array_equal(g["vertices"].view(np.uint32), want.view(np.uint32))
                  and g["num_internal"] == ni and np.array_equal(g["external_keys"], ref["keys"][ni:])
                  and np.array_equal(g["triangles"], ref["triangles"]))
            nv += len(want)
        identical = identical and bool(ok)
    return {"bins": len(picks), "identical": identical, "bin_indices": picks, "vertices_compared": nv,
            "against": "CPU oracle (oracle/), itself pinned against the reference's OpenCL kernels (tests/test_gpu_refcl.py)"}


def oracle_sample(orc, wl, indices, threads):
    """Runs the oracle over the given bins with `threads` host threads, one bin per thread, every stage
    (octree, MLS, marching) of a bin on that thread: the CPU's best use of its cores, since only the MLS
    stage of the oracle is parallel inside a bin.  Returns (splat instances, seconds)."""
    import ctypes
    from concurrent.futures import ThreadPoolExecutor
    gomp = ctypes.CDLL("libgomp.so.1")

    def work(i):
        gomp.omp_set_num_threads(1)          # per calling thread: no nested teams
        oracle_bin(orc, wl, wl["bins"][i])
        return wl["bins"][i]["n"]

    t0 = time.time()
    with ThreadPoolExecutor(max_workers=threads) as pool:
        splats = sum(pool.map(work, indices))    # ctypes releases the GIL inside the oracle
    return splats, time.time() - t0


def cpu_sample_plan(wl, threads, budget_s):
    """The sample of bins one measurement uses: the largest bins first (they are the typical ones),
    as many as fit the time budget, at least one per thread.  Returns (oracle, indices, calibration s)."""
    from oracle.oracle import Oracle
    orc = Oracle()
    orc.marching_tables()                        # build the static tables before the threads start
    order = sorted(range(len(wl["bins"])), key=lambda i: -wl["bins"][i]["n"])
    first = order[:min(threads, len(order))]
    _, dt1 = oracle_sample(orc, wl, first, threads)            # calibration: one bin per thread
    rounds = max(1, int(budget_s / max(dt1, 1e-3)))
    return orc, order[:min(len(order), rounds * len(first))], dt1


def cpu_baseline(wl, budget_s=15.0):
    threads = os.cpu_count() or 1
    orc, indices, _ = cpu_sample_plan(wl, threads, budget_s)
    splats, dt = oracle_sample(orc, wl, indices, threads)
    return splats / dt, len(indices), splats, dt, threads


CPU_SAMPLE_NOTE = "%d largest bins (%d splat instances) %s; %d host threads, one bin per thread through every stage of " \
                  "the oracle (octree, MLS, marching)"


def reference_opencl(wl, runner, nbins=3):
    """The 'reference OpenCL on one B200' column of BASELINE.json's metric: the reference's OWN
    processCorners kernel (kernels/mls.cl, embedded in oracle/_ref/librefcl.so) run by NVIDIA's OpenCL
    on this GPU over the largest bins, next to our kernel's CUDA-event time on the same bins.  Only the
    dominant kernel: the reference's host classes (Boost, clogs) cannot be built, see oracle/refcl/."""
    try:
        from oracle import refcl
        from oracle.oracle import Oracle
    except ImportError as e:
        return {"unavailable": str(e)}
    ref = refcl.get()
    if ref is None:
        return {"unavailable": refcl.why()}
    orc = Oracle()
    order = sorted(range(len(wl["bins"])), key=lambda i: -wl["bins"][i]["n"])[:nbins]
    w = runner.single
    w.set_profiling(1)
    w.reset_statistics()
    ref_s, corners, max_err = 0.0, 0, 0.0
    for i in order:
        b = wl["bins"][i]
        s = wl["splats"][b["first"]:b["first"] + b["n"]].copy()     # the build overwrites the radii: never on the workload itself
        tree = orc.octree_build(s, 0, len(s), b["offset"], 3, 6)
        size = [int(v) for v in b["size"]]
        pitch, zs, depth = (-(-v // 8) * 8 for v in size)
        img = np.zeros((zs * depth, pitch), np.float32)
        secs, _ = ref.mls_process_corners(img, s, tree["commands"][:tree["used"] + 1], tree["start"], 9, b["offset"], size[0],
                                          size[1], zs, 0, 0, size[2] - 1, orc.boundary_factor(1.0), 0)
        ref_s += secs
        corners += pitch * zs * depth
        w.upload_splats(np.ascontiguousarray(wl["splats"][b["first"]:b["first"] + b["n"]]))
        w.process_bin(0, b["n"], b["size"], b["offset"], collect=False)
    ours_s, _ = w.statistic(MLS_STAT)
    w.set_profiling(0)
    return {"kernel": "processCorners", "device": ref.device, "bins": len(order),
            "reference_ms": 1e3 * ref_s, "ours_ms": 1e3 * ours_s, "speedup": ref_s / ours_s if ours_s else None,
            "reference_corners_per_s": corners / ref_s if ref_s else None, "ours_corners_per_s": corners / ours_s if ours_s else None,
            "note": "the reference's own OpenCL kernel (kernels/mls.cl) through NVIDIA OpenCL on the same GPU, OpenCL profiling "
                    "events, against our CUDA kernel's events on the same %d largest bins" % len(order)}


def run_reference(args, rank, world):
    """--impl reference: the CPU port of the reference kernels (the reference itself cannot be
    built here: Boost/clogs/OpenCL/cppunit absent), all host threads, bounded sample per step."""
    if rank != 0:
        return
    wl = make_workload(args.splats, 2, 0, 1 if args.replicas else world)
    threads = os.cpu_count() or 1
    orc, indices, _ = cpu_sample_plan(wl, threads, 120.0 / max(1, args.steps + args.warmup))
    times, splats = [], 0
    for it in range(args.warmup + args.steps):
        ns, dt = oracle_sample(orc, wl, indices, threads)
        if it >= args.warmup:
            times.append(dt)
            splats = ns
    ms = 1e3 * float(np.mean(times))
    value = splats / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": "Msplats/s", "value": value, "unit": "Msplats/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, args),
            "cpu_baseline": {"value": value, "unit": "Msplats/s", "cores": threads, "kind": "port",
                             "sample": CPU_SAMPLE_NOTE % (len(indices), splats, "per step", threads)},
            "e2e": {"value": value, "unit": "Msplats/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(wl, args):
    if wl["world"] > 1:
        name = ("C2 x %d: ONE %d-splat synthetic bumpy torus, fit-grid %.3g, sphere fit, smooth %.0f, levels 6, subsampling 3, "
                "bins <=255 cells dealt to the %d ranks in contiguous angular runs of equal load (%d bins on rank 0)"
                % (wl["world"], wl["num_splats"], FIT_GRID / np.sqrt(wl["world"]), SMOOTH, wl["world"], len(wl["bins"])))
    else:
        name = ("C2: %d-splat synthetic bumpy torus, fit-grid %.3g, sphere fit, smooth %.0f, levels 6, "
                "subsampling 3, %d bins <=255 cells" % (wl["num_splats"], FIT_GRID, SMOOTH, len(wl["bins"])))
    return {"workload": name,
            "splats": wl["num_splats"], "bins": len(wl["bins"]), "splat_instances": int(len(wl["splats"])),
            "device_threads": args.threads, "bin_order": args.order,
            "readback": os.environ.get("MLSB_READBACK") or "copy",
            "l2": "inputs larger than L2: %.0f MB of splats and a 64 MiB field per bin cycle through a 126 MB L2"
                  % (len(wl["splats"]) * 32 / 1e6)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--splats", type=int, default=10_000_000)
    ap.add_argument("--threads", type=int, default=0,
                    help="workers per GPU (--device-threads); 0 = one per host core of this rank's share, between 3 and 12")
    ap.add_argument("--replicas", action="store_true",
                    help="N > 1: every rank meshes its own 10M torus (round-1 behaviour) instead of a share of one N x 10M torus")
    ap.add_argument("--order", default="largest-first", choices=["largest-first", "bucket"],
                    help="order in which the bins are handed to the worker group")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.threads <= 0:
        # one worker per host core of this rank's share of the box, between 3 and 12: on one B200 3 workers
        # gave 250, 4: 266, 8: 308, 12: 314, 16: 311 Msplats/s (round 1), and on the 32-vCPU 8-GPU box 4 workers
        # per GPU beat 12 (1315 against 1236 Msplats/s: oversubscribed host threads delay every sync)
        args.threads = max(3, min(12, (os.cpu_count() or 8) // world))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mlsgpu_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    numa_node = None
    if world > 1 and os.environ.get("MLSB_NUMA_BIND", "1") != "0":
        from mlsgpu_b200 import host as _host
        numa_node = _host.bind_to_gpu_numa_node(local)   # before any pinned allocation / worker thread
        log("rank %d: GPU %d on NUMA node %s, %d CPUs" % (rank, local, numa_node, len(os.sched_getaffinity(0))))
    # more spinning worker threads than cores would fight each other: block on events instead
    if world * args.threads > (os.cpu_count() or 1) * 3 // 4:
        os.environ.setdefault("MLSB_SYNC", "block")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG is left to the environment (the driver reads the communicator set-up from it); NCCL would
        # print to stdout, which carries the one JSON line, so its log goes to stderr unless a file is named
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import mlsgpu_b200 as m
    if args.replicas or world == 1:
        wl = make_workload(args.splats, 2 + rank)
        total_splats = args.splats * world
    else:
        wl = make_workload(args.splats, 2, rank, world)
        total_splats = wl["num_splats"]
    peak_inst = m.measure_fp32_peak(local)
    runner = Runner(wl, local, args.threads, args.order)

    def barrier():
        runner.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(from_host, steps, profile, single=False):
        for w in runner.all_workers():
            w.set_profiling(1 if profile else 0)
            w.reset_statistics()
        runner.mesh_bytes = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = sum(w.launch_count for w in runner.all_workers())
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            runner.step(from_host, single)
        runner.sync()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        launches = sum(w.launch_count for w in runner.all_workers()) - l0
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall * 1e3, launches

    # ---- work model (untimed): T pair tests and H hits of one step ----
    for w in runner.all_workers():
        w.set_work_counting(True)
        w.reset_statistics()
    runner.step(False)
    runner.sync()
    T = sum(w.mls_work()[0] for w in runner.all_workers())
    H = sum(w.mls_work()[1] for w in runner.all_workers())
    for w in runner.all_workers():
        w.set_work_counting(False)
    corners = sum(int(np.prod([-(-int(v) // 8) * 8 for v in b["size"]])) for b in wl["bins"])

    # ---- warm-up, then the two timed regions ----
    for _ in range(args.warmup):
        runner.step(False)
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, wall_dev, launches = timed(False, args.steps, profile=False)
    # one more step of the same kind, untimed, with events around the MLS launches: what the kernel's per-launch
    # times add up to while the workers' kernels overlap (a note in the roofline object)
    timed(False, 1, profile=True)
    mls_conc_s, _ = runner.stat_sum([MLS_STAT])
    mls_conc_s *= args.steps
    runner.step(True)   # warm the host path
    ms_e2e, wall_e2e, _ = timed(True, args.steps, profile=False)
    d2h = runner.mesh_bytes // args.steps
    readback = os.environ.get("MLSB_READBACK") or "copy"
    if readback == "compact":
        # what crossed PCIe is the wire blob (8 bytes per key, vertex and triangle), not the HostKeyMesh blob the
        # sink receives: the workers count it (statistic readback.bytes); same blocks in every pass, so the share
        # of the e2e steps is the share of all blocks since the statistics were reset
        wire, blocks = runner.stat_sum(["readback.bytes"])
        if blocks:
            d2h = int(d2h * wire / max(1.0, float(runner.blob_bytes_since_reset())))
    h2d = int(len(wl["splats"]) * 32)
    if world > 1 and not args.replicas:
        t = torch.tensor([float(d2h), float(h2d)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        d2h, h2d = int(t[0].item()), int(t[1].item())
    elif world > 1:
        d2h, h2d = d2h * world, h2d * world
    # roofline pass: the same steps on ONE worker, so that processCorners never shares the GPU with
    # another worker's kernels while its CUDA events are taken
    ms_single, _, _ = timed(False, args.steps, profile=True, single=True)
    mls_s, mls_n = runner.stat_sum([MLS_STAT])
    clocks = sampler.stop()

    # ---- per-stage breakdown: an untimed pass of the lone worker with every kernel timed, so that
    # the stage times are kernel times and not inflated by other workers' kernels on the same SMs ----
    for w in runner.all_workers():
        w.set_profiling(2)
        w.reset_statistics()
    runner.step(False, single=True)
    runner.sync()
    oct_s, _ = runner.stat_sum(OCTREE_STATS)
    mar_s, _ = runner.stat_sum(MARCHING_STATS)
    mls2_s, _ = runner.stat_sum([MLS_STAT])
    mc, mk, mv, mi, mw = runner.single.marching_work()
    if rank == 0:
        for name in OCTREE_STATS + [MLS_STAT] + MARCHING_STATS:
            s_, c_ = runner.stat_sum([name])
            log("  %-44s %9.3f ms/step  %5d launches" % (name, 1e3 * s_, c_))
    for w in runner.all_workers():
        w.set_profiling(0)
    parity = parity_check(runner, wl, int(time.time())) if rank == 0 else None

    value = total_splats * args.steps / (ms_dev * 1e-3) / 1e6
    e2e = total_splats * args.steps / (ms_e2e * 1e-3) / 1e6
    inst = 7.0 * T + 16.0 * H                        # FP32 pipe lane-instructions per step (this rank)
    mls_per_step = mls_s / args.steps
    achieved = 2.0 * inst / mls_per_step / 1e12      # TFLOP/s at 2 flop per lane-instruction
    peak = 2.0 * peak_inst / 1e12
    in_step = 2.0 * inst / (ms_dev * 1e-3 / args.steps) / 1e12
    inst_bytes = len(wl["splats"])
    # SURVEY 8d marching model: field read 4 B per corner; per occupied cell 24 W + 56 R; per unwelded (per-cell)
    # vertex 24 W, a 4-pass sort of (8 B key, 4 B index) pairs read + written, 36 R; per welded vertex 24 W;
    # per index 4 W + 12 (reindex)
    mar_bytes = 4.0 * mc + 80.0 * mk + (24.0 + 4 * 2 * 12 + 36.0) * mv + 24.0 * mw + 16.0 * mi
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", PROFILE_JSON)) as f:
            traffic = json.load(f)["processCorners"]["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    hbm_peak = 6541.1
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak = float(json.load(f)["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    line = {
        "metric": "Msplats/s", "value": value, "unit": "Msplats/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(wl, args),
        "e2e": {"value": e2e, "unit": "Msplats/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity_checked": parity,
        "mls_corners_per_s": corners * args.steps / mls_s if mls_s > 0 else None,
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_note": "DRAM bytes per launch on a large bin, ncu --set full "
                     "(profiles/%s); the kernel is FP32-issue-bound, not HBM-bound" % PROFILE_JSON,
                     "kernel": "processCornersKernel",
                     "model": "7*T + 16*H FP32 pipe instructions per step (SURVEY 8d), 2 flop each; T counts the pair tests of "
                              "the reference algorithm, most of which the warp cull never executes",
                     "pair_tests": int(T), "hits": int(H), "kernel_ms_per_step": 1e3 * mls_per_step,
                     "launches_per_step": mls_n // args.steps,
                     "in_step": {"frac": in_step / peak, "achieved": in_step,
                                 "note": "the same model work over the WHOLE timed step of the %d-worker group (octree, marching "
                                         "and copies of other bins share the GPU): a lower bound on what the kernel sustains there; "
                                         "its per-launch event times in that region add up to %.1f ms/step because the workers' "
                                         "kernels overlap" % (args.threads, 1e3 * mls_conc_s / args.steps)},
                     "measured_in": "timed single-worker pass (%d steps, %.1f ms/step): the kernel alone on the GPU, CUDA events "
                                    "on its stream" % (args.steps, ms_single / args.steps),
                     "peak_source": "measured here: dependent-free FFMA stream, %.2f T lane-inst/s "
                                    "(MEASURED_PEAKS.json has no FP32 figure)" % (peak_inst / 1e12)},
        "stages_ms_per_step": {"octree": 1e3 * oct_s, "mls": 1e3 * mls2_s, "marching": 1e3 * mar_s,
                               "wall": wall_dev / args.steps,
                               "note": "kernel times of one worker running alone (CUDA events around every launch); "
                                       "wall = the timed step of the worker group"},
        "rooflines_hbm": {
            "octree": {"bound": "hbm", "model_bytes": 472 * inst_bytes, "achieved": 472 * inst_bytes / oct_s / 1e9 if oct_s else None,
                       "peak": hbm_peak, "unit": "GB/s", "frac": 472 * inst_bytes / oct_s / 1e9 / hbm_peak if oct_s else None,
                       "model": "472 B per splat instance (SURVEY 8d: the reference's 8-slot entries, 2-pass pair sort, scan, emit)",
                       "traffic": "profiles/%s" % PROFILE_JSON},
            "marching": {"bound": "hbm", "model_bytes": mar_bytes, "achieved": mar_bytes / mar_s / 1e9 if mar_s else None,
                         "peak": hbm_peak, "unit": "GB/s", "frac": mar_bytes / mar_s / 1e9 / hbm_peak if mar_s else None,
                         "model": "SURVEY 8d: 4 B x %d corners + 80 B x %d occupied cells + 156 B x %d per-cell vertices "
                                  "+ 24 B x %d welded vertices + 16 B x %d indices" % (mc, mk, mv, mw, mi),
                         "traffic": "profiles/%s" % PROFILE_JSON},
            "note": "stage kernel times of the lone worker; C2's bins hold ~150 k splats in 256^3 corners, so both stages are "
                    "launch-latency-bound there (tools/bench_stages.py measures a 2 M-splat bin)"},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, done, ns, dt, threads = cpu_baseline(wl)
        line["cpu_baseline"] = {"value": rate / 1e6, "unit": "Msplats/s", "cores": threads, "kind": "port",
                                "sample": CPU_SAMPLE_NOTE % (done, ns, "in %.1f s" % dt, threads)}
        line["reference_opencl_b200"] = reference_opencl(wl, runner)
    if rank == 0:
        print(json.dumps(line), flush=True)
    runner.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
