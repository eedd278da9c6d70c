#!/usr/bin/env python
"""Benchmark of the discretization + flux hot path (BASELINE.json metric: WE frames/s assigned and
flux-accumulated; 1 frame = 1 WE segment in 1 iteration = 2 feature vectors assigned + 1 weighted
transition scattered).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5|cfg2|cfg3|cfg4|...] [--impl b200|reference]

Default workload = BASELINE config 5's per-frame shape (8,000 segs x 256-dim features, 100 WE bins x 100 clusters),
the largest configuration whose iterations can be held resident on ONE B200: 4,000 of its 5,000 iterations (131 GB
of features; the full 164 GB does not fit beside the workspaces).  One step = one "haMSM rebuild" pass over the
resident iterations, as config 5 words it:
    10 Lloyd iterations  (K0 bins of the parent pcoords once; per iteration K1 assignment of the child frames +
                          K2 order-deterministic centroid accumulation + mean; centres reset at the start of a step)
  + assignment of every parent and child frame against the refined centres (K0 + K1)
  + flux accumulation of every transition (K3 sort + segmented fp64 sum into the dense matrix, / nI).
With N > 1 ranks the SAME iterations are split by contiguous iteration range (strong scaling); the exchange steps
are the all-reduce of the Lloyd partial sums (20.6 MB per Lloyd iteration) and of the flux matrix.

value  : frames/s of the whole job, inputs resident in HBM (CUDA events per step, inputs >> L2 and L2 flushed
         between steps, max over ranks);
e2e    : the same pass through the public API (modelWE.lloyd_refine_clusters + launch_ray_discretization +
         get_fluxMatrix) on HOST numpy buffers: every H2D / D2H copy is inside the timed region (each frame crosses
         PCIe once per pass: the refinement leaves the rows it shipped on the device for the discretization);
roofline: the dominant kernel (K1 of the final assignment) AND the whole step, against the measured HBM copy
         bandwidth and the cuBLAS DGEMM rate measured on this box;
cpu_baseline / --impl reference: the reference's own CPU pattern on the host cores (sklearn KMeans Lloyd per WE
         bin, one sklearn predict([x]) per segment, scipy coo_matrix -> dense add per iteration), one process per
         iteration range, on a bounded sample of the same workload;
extra  : short secondary measurements (config 2, config 3 shape, config 4 flux stress).
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "we_frames_per_sec_assigned_and_flux_accumulated"
UNIT = "frames/s"

# iterations of each named shape held resident by ONE GPU (N > 1 splits the same iterations by range)
RESIDENT_ITERS = {"cfg5": 4000, "cfg3": 300, "cfg2": 200, "cfg5s": 250, "cfg3s": 50, "tiny": 12}
LLOYD_ITERS = {"cfg5": 10, "cfg5s": 10}


_T0 = time.time()


def log(msg):
    """Progress on stderr (rank 0 only by convention of the callers): a stuck leg must be identifiable from the log."""
    print(f"[bench {time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--iters", type=int, default=0, help="WE iterations of the shape held resident in total (0 = table)")
    ap.add_argument("--lloyd-iters", type=int, default=-1, help="Lloyd iterations per step (-1 = the workload's definition)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-iters", type=int, default=0, help="iterations of the workload timed on the CPU (0 = auto)")
    ap.add_argument("--precision-path", default="auto", choices=["auto", "fp64", "tf32x3"],
                    help="K1 evaluation: fp64 DMMA, or tcgen05 split-TF32 candidates + fp64 re-check (same labels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--e2e-iters", type=int, default=0, help="iterations of the shape the e2e leg holds on the host (0 = auto)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_text(cfg, iters, lloyd):
    step = (f"{lloyd} Lloyd iterations (K1 assign of child frames + K2 centroid update) + " if lloyd else "") + \
        "assign of parent+child frames (K0+K1) + flux accumulation (K3, / nI)"
    return (f"{cfg.name} shape: {iters} of {cfg.n_iters} WE iters x {cfg.n_segs} segs x {cfg.dim}-dim, {cfg.n_bins} bins x "
            f"{cfg.k_per_bin} clusters/bin, fp64; step = {step}")


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's literal pattern, one process per iteration range.  Imports only numpy / sklearn / scipy,
# workloads.py and oracle/ -- never the msm_we_b200 package (the reference arm must not map the CUDA library).
# ----------------------------------------------------------------------------------------------
_POOL = {}
_CPU_CACHE = {}
_CPU_ARGS = {}


def _close_pools():
    for pool in _POOL.values():
        pool.close()
        pool.join()
    _POOL.clear()


def _cpu_inputs(cfg_name, n_iters):
    key = (cfg_name, n_iters)
    if key not in _CPU_CACHE:
        import workloads

        cfg = dataclasses.replace(workloads.CONFIGS[cfg_name], n_iters=n_iters)
        means, centers = workloads.make_centers(cfg)
        its = workloads.generate_host(cfg, means)
        _CPU_CACHE[key] = (cfg, centers, its)
    return _CPU_CACHE[key]


def _cpu_chunk(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    import workloads
    from oracle import oracle as O

    cfg_name, lo, hi = args
    n_sample, lloyd = _CPU_ARGS["sample"], _CPU_ARGS["lloyd"]
    cfg, centers, its = _cpu_inputs(cfg_name, n_sample)
    basis, target = workloads.region_bounds(cfg)
    om = O.RectilinearBinMapperOracle(workloads.boundaries(cfg))
    if lloyd > 0:
        # reference: KMeans.fit -> lloyd_iter_chunked_dense (msm_we/_hamsm/_clustering.py:289,491), one model per WE bin
        # on this range's child frames, binned by the PARENT pcoord with basis/target parents dropped (:849-877)
        from sklearn.cluster import KMeans

        X = np.concatenate([its[i]["child"] for i in range(lo, hi)])
        pc0 = np.concatenate([its[i]["pcoord0"] for i in range(lo, hi)])
        keep = ~(O.is_we_region(pc0, basis) | O.is_we_region(pc0, target))
        bins = om.assign(pc0[keep])
        Xk = X[keep]
        refined = []
        for b in range(cfg.n_bins):
            rows = Xk[bins == b]
            if rows.shape[0] >= cfg.k_per_bin:
                km = KMeans(n_clusters=cfg.k_per_bin, init=centers[b], n_init=1, max_iter=lloyd, tol=0.0, algorithm="lloyd")
                km.fit(rows)
                refined.append(np.ascontiguousarray(km.cluster_centers_))
            else:
                refined.append(centers[b])
        centers = refined
    strat = O.StratifiedOracle(om, centers, basis, target)
    models = [O.make_fitted_minibatch(c) for c in centers]
    n = cfg.n_clusters
    total = np.zeros((n + 2, n + 2))
    frames = 0
    for i in range(lo, hi):
        d = its[i]
        parent, child = O.discretize_iteration(strat, d["parent"], d["child"], d["pcoord0"], d["pcoord1"], literal=True,
                                               models=models)
        pairs = np.stack([parent, child], axis=1)
        total = total + O.iter_flux_matrix(n, pairs, d["pcoord0"], d["pcoord1"], d["weights"], basis, target)
        frames += len(parent)
    return frames, float(total.sum())


def cpu_reference_pass(cfg_name, n_sample_iters, n_procs, lloyd):
    """Times the reference pattern on `n_sample_iters` iterations of the workload; returns (frames, seconds, cores).
    The worker pool is created once (fork) and reused by every step."""
    import multiprocessing as mp

    _CPU_ARGS.update(sample=n_sample_iters, lloyd=lloyd)
    _cpu_inputs(cfg_name, n_sample_iters)           # build before forking
    import sklearn.cluster  # noqa: F401  (imported in the parent so the forked workers inherit it)
    from oracle import oracle as O

    O.make_fitted_minibatch(np.zeros((2, 2))).predict([[0.0, 0.0]])
    chunks = [c for c in np.array_split(np.arange(n_sample_iters), n_procs) if len(c)]
    key = (cfg_name, n_sample_iters, len(chunks), lloyd)
    if key not in _POOL:
        _POOL[key] = mp.get_context("fork").Pool(len(chunks))
    t0 = time.perf_counter()
    res = _POOL[key].map(_cpu_chunk, [(cfg_name, int(c[0]), int(c[-1]) + 1) for c in chunks], chunksize=1)
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt, len(chunks)


def cpu_plan(cfg, cores, n_steps, budget_s):
    """(processes, iterations per step): every process holds two dense (n+2)^2 matrices (the reference's per-iteration
    `.todense()` and the running sum), which bounds the process count by host memory; the whole warm-up + timed run is
    sized to about `budget_s` seconds, at least one iteration per process."""
    M = cfg.n_clusters + 2
    try:
        ram = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        ram = 64 << 30
    procs = int(max(1, min(cores, (0.5 * ram) // (3 * M * M * 8 + (1 << 30)))))
    per_iter = cfg.n_segs * 2 * (150e-6 + 2.5e-10 * cfg.k_per_bin * cfg.dim) + 6e-9 * M * M   # core-seconds, measured shape
    per_proc = max(1, int(budget_s / max(n_steps, 1) / per_iter))
    return procs, int(min(cfg.n_iters, per_proc * procs))


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled by an `nvidia-smi -lms` subprocess started before the warm-up and
    stopped after the timed region (B200_PROFILING.md recipe); samples are split by wall-clock time into
    the timed window and the rest."""

    QUERY = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = os.path.join("/tmp", f"mwe_clocks_{os.getpid()}_{index}.csv")
        self.t_on = self.t_off = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20", "-f", self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            t_end = time.time() + 5.0
            while time.time() < t_end:          # nvidia-smi needs up to a second to start on a fresh box
                try:
                    if os.path.getsize(self.path) > 0:
                        break
                except OSError:
                    pass
                time.sleep(0.02)
        except Exception:
            self.proc = None
        return self

    def mark_timed(self, on):
        if on:
            self.t_on = time.time()
        else:
            self.t_off = time.time()

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        import datetime

        rows = []
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [x.strip() for x in line.split(",")]
                    if len(parts) >= 7:
                        try:
                            ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                            rows.append((ts, float(parts[1]), float(parts[2]), parts[3:7]))
                        except ValueError:
                            pass
            os.remove(self.path)
        except OSError:
            pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        timed = [r for r in rows if self.t_on is not None and self.t_on - 0.02 <= r[0] <= (self.t_off or 1e18) + 0.02]
        use = timed if timed else rows
        sm = sorted(r[1] for r in use)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3][i].lower().startswith("active") for r in use)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": use[0][2], "reasons": reasons, "samples": len(use),
                "window": "timed region" if timed else "warm-up + timed region (timed region shorter than the sampling period)"}


def measure_dgemm(dev):
    """cuBLAS DGEMM rate on this box (the fp64 tensor-roofline denominator): torch.matmul float64 4096^3, best of 5."""
    import torch

    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def k3_passes(M, C=1):
    bits = int(np.ceil(np.log2(float(C * M) ** 2 + 1)))
    return (bits + 7) // 8


def launches_per_step(cfg, tc_path, lloyd, world):
    """Kernels of ours per step, counted from the launch sequences in csrc/ (confirmed by the ncu launch lists in
    profiles/)."""
    k1 = 2 + (3 if tc_path else 0) + 2                      # scan, scatter, [mean, split, csq], main, re-check
    k2 = 2 + 1 + max(1, (int(np.ceil(np.log2(cfg.n_clusters + 1))) + 7) // 8) + 1 + 1  # keys, bounds, hist + passes, order, sum
    per_lloyd = 1 + (k1 - 2) + k2 + 1                      # csq + K1 without bucketing + K2 + finalize (relocation kernels not counted)
    first_lloyd = 3                                        # count, scan, scatter: the points are bucketed by WE bin once per fit
    final = 1 + k1 + 1 + (1 + k3_passes(cfg.n_clusters + 2)) + 3 + 1      # K0 (counts the buckets), K1, keys, sort, mark/group/cell, divide|exchange
    return (1 if lloyd else 0) + lloyd * per_lloyd + (first_lloyd if lloyd else 0) + (1 if lloyd else 0) + final


def step_work(cfg, n_frames, lloyd, P=1):
    """Algorithmic bytes and fp64 FLOPs of one step (SURVEY section 8d per-frame figures x frames)."""
    D, K = cfg.dim, cfg.k_per_bin
    assign_flux = 2 * D * 8 + 2 * P * 8 + 8 + 2 * 8                 # fused assign+flux row: read + label write
    lloyd_iter = (D * 8 + P * 8 + 8) + (D * 8 + 8 + 8)              # K1 on the child frame + K2
    bytes_ = n_frames * (assign_flux + lloyd * lloyd_iter)
    flops = n_frames * (4 * K * D + 1 + lloyd * (2 * K * D + 2 * D))
    return bytes_, flops


# ----------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        os.environ["OMP_NUM_THREADS"] = "1"        # the reference requires it (msm_we/msm_we.py:75-80); set before sklearn loads
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import workloads

    if args.workload == "cfg4":
        return main_cfg4(args, rank, world, local_rank)
    cfg = workloads.CONFIGS[args.workload]
    iters_total = args.iters or RESIDENT_ITERS.get(args.workload, cfg.n_iters)
    lloyd = LLOYD_ITERS.get(args.workload, 0) if args.lloyd_iters < 0 else args.lloyd_iters
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        procs, sample = cpu_plan(cfg, cores, args.steps + args.warmup, budget_s=150.0)
        log(f"reference arm: {procs} processes, {args.cpu_sample_iters or sample} iterations per step")
        sample = args.cpu_sample_iters or sample
        times, frames, used = [], 0, procs
        for step in range(args.warmup + args.steps):
            frames, dt, used = cpu_reference_pass(args.workload, sample, procs, lloyd)
            log(f"reference arm step {step}: {dt:.1f} s")
            if step >= args.warmup:
                times.append(dt)
        total = sum(times)
        val = frames * len(times) / total
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(cfg, iters_total, lloyd)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port",
                             "sample": f"{sample} iterations of the shape per step ({frames} frames): sklearn KMeans Lloyd per "
                                       f"WE bin, literal reference loop (sklearn predict([x]) per segment + scipy coo->dense "
                                       f"add per iteration), one process per iteration range (ray unavailable: multiprocessing)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; the B200 arm has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if rank == 0:
        log(f"b200 arm: {workload_text(cfg, iters_total, lloyd)}; world {world}")
    res = run_b200(cfg, args.workload, iters_total, lloyd, args.precision_path, args.steps, args.warmup, rank, world, dev,
                   local_rank)
    if rank == 0:
        log(f"resident measurement done: {res['ms_per_step']:.3f} ms/step, {res['value']:.4g} frames/s")
    line = None
    if rank == 0:
        line = res
    # ---------------------------------------------------------------- e2e through the public API
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(cfg, lloyd, rank, world, dev, args.e2e_iters)
        if rank == 0:
            log(f"e2e done: {e2e['value']:.4g} frames/s (staged {e2e['staged_value']:.4g}, first call {e2e['first_call_value']:.4g})")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_baseline(args, cfg, cores)
    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        extra = run_extra(dev, local_rank, args.workload)
        log("extra lines done")
    if rank == 0:
        line["e2e"] = e2e
        line["cpu_baseline"] = cpu
        line["extra"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_cpu_baseline(args, cfg, cores):
    """The CPU leg runs in a FRESH interpreter (this file with --impl reference, one step): the worker processes are
    forked, and forking a process that has initialised CUDA, page-locked gigabytes and started OpenMP / staging thread
    pools is where fork-unsafe libraries deadlock (a forked sklearn KMeans waits forever on the parent's OpenMP pool)."""
    procs, sample = cpu_plan(cfg, cores, 1, budget_s=20.0)
    sample = args.cpu_sample_iters or sample
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "1",
           "--warmup", "0", "--cpu-sample-iters", str(sample)]
    if args.iters:
        cmd += ["--iters", str(args.iters)]
    if args.lloyd_iters >= 0:
        cmd += ["--lloyd-iters", str(args.lloyd_iters)]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    log(f"cpu_baseline: {' '.join(cmd[2:])}")
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env)
        line = json.loads(out.stdout.strip().splitlines()[-1])
        cb = line["cpu_baseline"]
        cb["sample"] += f", {line['ms_per_step'] / 1e3:.1f} s wall"
        return cb
    except Exception as e:
        log(f"cpu_baseline failed: {e!r}")
        return {"value": None, "unit": UNIT, "cores": procs, "kind": "port", "sample": f"failed: {e!r}"[:300]}


def run_b200(cfg, name, iters_total, lloyd, precision_path, steps, warmup, rank, world, dev, local_rank, quiet_clocks=False):
    """Resident-input measurement of one workload; returns the JSON line (dict) on every rank."""
    import torch
    import torch.distributed as dist

    import workloads
    from msm_we_b200 import _lib, clustering_ops, ops
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.engine import DeviceClusters

    # this rank's contiguous iteration range of the SAME total (strong scaling)
    lo, hi = iters_total * rank // world, iters_total * (rank + 1) // world
    my_iters = hi - lo
    means, centers = workloads.make_centers(cfg)
    basis, target = workloads.region_bounds(cfg)
    mapper = RectilinearBinMapper(workloads.boundaries(cfg))
    engine = DeviceClusters(mapper, centers, {b: b for b in range(cfg.n_bins)}, basis, target, 1, device=dev)
    centers0 = engine.centers.clone()
    data = workloads.generate_device(cfg, dev, means=means, seed_offset=rank, iters=my_iters)
    torch.cuda.synchronize()
    if rank == 0:
        log(f"{name}: {my_iters} iterations generated on the device ({data['X'].numel() * 8 / 1e9:.1f} GB of features)")
    N = data["n"]
    n_clusters = cfg.n_clusters
    M = n_clusters + 2
    X, pc, w, offs = data["X"], data["pcoord"], data["weights"], data["iter_offsets"]
    sumK = engine.centers.shape[0]
    group = dist.group.WORLD if world > 1 else None
    # N > 1: small matrices go through the one-kernel peer-memory exchange, large ones through NCCL
    reducer = None
    if world > 1 and M * M * 8 <= (64 << 20):
        from msm_we_b200.distributed import PeerFluxAllreduce
        reducer = PeerFluxAllreduce.create((M, M), dev)
    # N > 1 with a large matrix: the NCCL all-reduce (+ / nI) of step k runs on a side stream while step k+1 starts its
    # Lloyd iterations (the reduced matrix is only needed at the end of the pass); two matrices alternate
    overlap = world > 1 and reducer is None
    n_dense = 2 if overlap else 1
    denses = [reducer.partial] if reducer is not None else [torch.zeros((M, M), dtype=torch.float64, device=dev)
                                                            for _ in range(n_dense)]
    side = torch.cuda.Stream(device=dev) if overlap else None
    flux_group = dist.new_group() if overlap else None      # its own communicator: runs beside the Lloyd all-reduces
    side_done = [None] * n_dense
    labels = torch.empty(2 * N, dtype=torch.int64, device=dev)
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    path = {"auto": _lib.ASSIGN_AUTO, "tf32x3": _lib.ASSIGN_TF32X3, "fp64": _lib.ASSIGN_FP64}[precision_path]
    if path == _lib.ASSIGN_AUTO:   # same rule as the library (csrc/assign.cu, resolve_assign_path)
        path = _lib.ASSIGN_TF32X3 if cfg.k_per_bin * cfg.dim >= 2048 else _lib.ASSIGN_FP64
    Xc, pc0 = X[N:], pc[:N]
    step_no = [0]

    def step(ev=None):
        k = step_no[0] % n_dense
        step_no[0] += 1
        dense = denses[k]
        if side_done[k] is not None:
            torch.cuda.current_stream().wait_event(side_done[k])      # the exchange that last used this matrix has finished
        dense.zero_()
        if lloyd:
            engine.centers.copy_(centers0)
            bins_p, flags_p = engine.bins_and_flags(pc0)
            clustering_ops.lloyd_fit(Xc, None, bins_p, engine.centers, engine.bin_offset, engine.max_k, lloyd, group=group,
                                     flags_dev=flags_p, path=path, errors=engine.errors)
            engine.csq = ops.centers_sqnorm(engine.centers)
        if ev is not None:
            _lib.set_timing_events(ev[0], ev[1])
        # one C call enqueues K0 -> K1 -> K3 (-> / nI when single-GPU)
        engine.hotpath_step(X, pc, w, n_clusters, iter_offsets=offs, dense=dense,
                            divisor=float(iters_total) if world == 1 else 0.0, labels_out=labels, path=path)
        if ev is not None:
            _lib.set_timing_events(None, None)
        if reducer is not None:
            reducer.reduce(float(iters_total))
        elif world > 1:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(side):
                side.wait_event(ready)
                dist.all_reduce(dense, group=flux_group)
                ops.divide_(dense, float(iters_total))
                done = torch.cuda.Event()
                done.record()
            side_done[k] = done

    def join():
        """Everything the steps put on the side stream is complete when the main stream passes this point."""
        for d in side_done:
            if d is not None:
                torch.cuda.current_stream().wait_event(d)

    clocks = ClockSampler(local_rank)
    if not quiet_clocks:
        clocks.__enter__()
    for _ in range(warmup):
        l2_flush.zero_()
        step()
    join()
    torch.cuda.synchronize()
    engine.check_errors()
    if rank == 0:
        log(f"{name}: warm-up done")

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.mark_timed(True)
    for k in range(steps):
        l2_flush.zero_()                       # flush L2 between timed steps (outside the events unless steps overlap)
        evs[k][0].record()
        step(kevs[k])
        if k == steps - 1:
            join()                             # the last step's exchange is inside the timed region
        evs[k][1].record()
    torch.cuda.synchronize()
    clocks.mark_timed(False)
    if not quiet_clocks:
        clocks.__exit__()
    if world > 1:
        dist.barrier()
    # steps overlap when the exchange runs on the side stream: then the time is first start -> last end (which includes
    # the L2 flushes between steps); otherwise the sum of the per-step intervals
    total_ms = evs[0][0].elapsed_time(evs[-1][1]) if overlap else sum(a.elapsed_time(b) for a, b in evs)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kevs) / steps
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    engine.check_errors()
    if reducer is not None:
        reducer.errors.check()
    frames_per_step = cfg.n_segs * iters_total
    value = frames_per_step * steps / (total_ms * 1e-3)
    ms_per_step = total_ms / steps

    # roofline: dominant kernel (K1 main kernel of the final assignment; 2N points per launch, per point D*8 feature
    # bytes + 4 (bucket index) + 8 (int64 label)) and the whole step
    peak, peak_src = load_peaks()
    dgemm = measure_dgemm(dev) if rank == 0 else None
    alg_bytes = 2 * N * (cfg.dim * 8 + 4 + 8)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    tc = path == _lib.ASSIGN_TF32X3
    n_pad = (cfg.k_per_bin + 15) // 16 * 16
    kname = ("assign_tc2_kernel" if 2 * n_pad + 128 <= 512 else "assign_tc_kernel") if tc else ("assign_dmma_resident_kernel" if (cfg.k_per_bin <= 64 and cfg.dim % 2 == 0)
                                           else "assign_dmma_kernel")
    step_bytes, step_flops = step_work(cfg, N, lloyd)        # this rank's share; ranks run concurrently
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_fp64_equiv_tflops": 2 * N * 2.0 * cfg.k_per_bin * cfg.dim / (kernel_ms * 1e-3) / 1e12,
                "step": {"algorithmic_bytes_per_gpu": step_bytes, "fp64_flops_per_gpu": step_flops,
                         "achieved_gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
                         "frac_hbm": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                         "achieved_fp64_tflops": step_flops / (ms_per_step * 1e-3) / 1e12,
                         "dgemm_tflops_measured": dgemm,
                         "frac_fp64_tensor": None if not dgemm else step_flops / (ms_per_step * 1e-3) / 1e12 / dgemm}}
    try:   # DRAM bytes per launch of that kernel from the committed ncu capture of the same shape (None if none)
        with open(os.path.join(ROOT, "profiles", "k1_dram_traffic.json")) as f:
            entry = json.load(f).get(f"{name}/{kname}")
        if entry:
            roofline["traffic"] = entry["bytes"] * (2 * N) / (entry.get("points") or 2 * N)
            roofline["traffic_source"] = entry["source"]
    except (OSError, ValueError):
        pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(cfg, iters_total, lloyd), "frames_per_step": frames_per_step,
                   "iterations_per_gpu": my_iters, "resident_feature_bytes_per_gpu": int(X.numel() * 8),
                   "l2": "inputs far larger than L2 and L2 flushed between timed steps (256 MiB memset)",
                   "exchange": ("none (1 GPU)" if world == 1 else
                                ("Lloyd partial sums: NCCL all-reduce per Lloyd iteration; " if lloyd else "") +
                                ("flux: one peer-memory kernel per rank (rank-order sum + / nI over NVLink)" if reducer is not None
                                 else "flux: NCCL all-reduce of the dense matrix + divide on a side stream, overlapping the next "
                                      "step's Lloyd iterations (two matrices alternate; the last step's exchange is inside the "
                                      "timed region)")),
                   "precision_path": ("tcgen05 split-TF32 candidate pass + fp64 re-check of near-ties (labels identical "
                                      "to the fp64 path)") if tc else "fp64 DMMA"},
        "roofline": roofline, "clocks": None if quiet_clocks else clocks.summary(),
        "gpu_launches": launches_per_step(cfg, tc, lloyd, world) * steps,
    }
    if reducer is not None:
        dist.barrier()
        reducer.close()
    del data, X, pc, w, labels, denses, l2_flush
    torch.cuda.empty_cache()
    return line


def run_e2e(cfg, lloyd, rank, world, dev, e2e_iters=0):
    """frames/s through the public modelWE API with HOST numpy buffers: every feature / pcoord / weight crosses PCIe
    inside the timed region, labels and the flux matrix come back.  Three figures: steady state with the model's own
    arrays page-locked in place (`value`), the very first pass (pays the one-time cudaHostRegister), and the staged
    path a data source that produces fresh arrays every pass takes (featuriser output, HDF5 reads)."""
    import torch
    import torch.distributed as dist

    import workloads
    from msm_we_b200 import _pinning
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.msm_we import modelWE
    from msm_we_b200.stratified_clustering import StratifiedClusters

    # a bounded number of iterations of the same shape (host generation is slow and host RAM is finite)
    n_it = e2e_iters or int(max(8, min(cfg.n_iters, (3 << 30) // (cfg.n_segs * 2 * cfg.dim * 8))))
    hcfg = dataclasses.replace(cfg, n_iters=n_it, seed=cfg.seed + rank)
    means, centers = workloads.make_centers(cfg)
    its = workloads.generate_host(hcfg, means)
    basis, target = workloads.region_bounds(cfg)

    def build():
        model = modelWE()
        model.initialize(workloads.to_iteration_source(its), None, "bench", basis_pcoord_bounds=basis,
                         target_pcoord_bounds=target, tau=1.0, pcoord_ndim=1)
        model.get_iterations()
        model.dimReduce()
        clusters = StratifiedClusters(RectilinearBinMapper(workloads.boundaries(cfg)), model, cfg.k_per_bin, [])
        for b in range(cfg.n_bins):
            clusters.cluster_models[b].cluster_centers_ = centers[b].copy()
        model.clusters = clusters
        model.n_clusters = cfg.n_clusters
        return model

    def once(model):
        if lloyd:
            for b in range(cfg.n_bins):
                model.clusters.cluster_models[b].cluster_centers_ = centers[b].copy()
            model.lloyd_refine_clusters(lloyd)
        model.launch_ray_discretization()
        model.get_fluxMatrix(n_lag=0, first_iter=0)
        return model.fluxMatrixRaw

    def timed(model, steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            once(model)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    frames = (n_it - 1) * cfg.n_segs          # launch_ray_discretization covers range(1, maxIter)
    model = build()
    first = timed(model, 1)                   # cold: page-locks the model's arrays, builds the device state
    steady = timed(model, 3)
    enabled = _pinning.PINS.enabled
    _pinning.PINS.enabled = False             # fresh arrays every pass: everything goes through the pinned staging rows
    try:
        staged_model = build()
        timed(staged_model, 1)
        staged = timed(staged_model, 2)
    finally:
        _pinning.PINS.enabled = enabled
    M = cfg.n_clusters + 2
    # features of parent and child frames + pcoords (the child rows cross once: lloyd_refine_clusters leaves them on the
    # device for the discretization that follows, single use), flux inputs, the Lloyd pass's own parent pcoords
    h2d = frames * (2 * (cfg.dim + 1) * 8) + frames * (2 * 8 + 8 + 2 * 8) + (frames * 8 if lloyd else 0)
    d2h = frames * 2 * (8 + 4 + 1) + M * M * 8 + (cfg.n_clusters * cfg.dim * 8 if lloyd else 0)
    return {"value": frames * world * 3 / steady, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "frames_per_step": frames * world,
            "first_call_value": frames * world / first, "staged_value": frames * world * 2 / staged,
            "api": ("modelWE.lloyd_refine_clusters + " if lloyd else "") + "modelWE.launch_ray_discretization + get_fluxMatrix, "
                   "numpy in / numpy out, each frame crosses PCIe once per pass; value = steady state (source arrays page-locked in place), first_call_value = cold "
                   "pass incl. cudaHostRegister, staged_value = every array through pinned staging rows"}


def run_extra(dev, local_rank, main_name):
    """Short secondary measurements (device-resident inputs, CUDA events): the other BASELINE shapes."""
    import workloads

    out = []
    for name, steps, warm in (("cfg2", 30, 5), ("cfg3", 3, 2)):
        if name == main_name:
            continue
        try:
            cfg = workloads.CONFIGS[name]
            iters = RESIDENT_ITERS[name] if name != "cfg3" else 50
            r = run_b200(cfg, name, iters, 0, "auto", steps, warm, 0, 1, dev, local_rank, quiet_clocks=True)
            out.append({"workload": r["config"]["workload"], "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                        "kernel": r["roofline"]["kernel"], "kernel_frac_hbm": r["roofline"]["frac"],
                        "step_frac_hbm": r["roofline"]["step"]["frac_hbm"], "steps": steps})
        except Exception as e:        # secondary lines never take the headline down
            out.append({"workload": name, "error": repr(e)[:200]})
    try:
        out.append(bench_cfg4(dev, n_transitions=1 << 26, steps=3, warmup=2))
    except Exception as e:
        out.append({"workload": "cfg4", "error": repr(e)[:200]})
    return out


def bench_cfg4(dev, n_transitions, steps, warmup, seed_offset=0):
    """BASELINE config 4 (flux-matrix stress): 20,000 clusters x 2 history colours, weighted transitions with label
    locality, K3 with C = 2 and sorted-COO output.  1e9 transitions are the 8-GPU total; one GPU takes its 1/8
    (1.25e8) or the smaller `n_transitions` given."""
    import torch

    import workloads
    from msm_we_b200 import ops

    d = workloads.generate_cfg4_device(dev, n_transitions, seed_offset=seed_offset)
    n, N = d["n_clusters"], d["n"]

    def step():
        return ops.flux_accumulate(d["start"], d["end"], d["w"], n, col0=d["col0"], col1=d["col1"], C=2,
                                   iter_offsets=d["iter_offsets"], want_coo=True, dense=None)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record(); res = step(); b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    nnz = int(res[1][3].item())
    peak, _ = load_peaks()
    alg = N * 26 + nnz * 24
    return {"workload": f"cfg4 flux stress: {n} clusters x 2 colours (M = {2 * (n + 2)}), {N} weighted transitions on this GPU "
                        f"(1/8 of 1e9 = 1.25e8), sorted-COO output", "value": N / (ms * 1e-3), "unit": "transitions/s",
            "ms_per_step": ms, "nnz": nnz, "algorithmic_bytes": alg, "step_frac_hbm": alg / (ms * 1e-3) / 1e9 / peak,
            "steps": steps}


def main_cfg4(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "cfg4 is a K3-only stress line; use the default workload"}))
        return 0
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.iters or (125_000_000 if world > 1 else 1 << 27)
    r = bench_cfg4(dev, n, args.steps, args.warmup, seed_offset=rank)
    if world > 1:
        t = torch.tensor([r["ms_per_step"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r["ms_per_step"] = float(t.item())
        r["value"] = n * world / (r["ms_per_step"] * 1e-3)
    if rank == 0:
        r.update(metric="weighted_transitions_per_sec_flux_accumulated", n_gpus=world, scaling="weak", higher_is_better=True)
        print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    try:
        rc = main()
    finally:
        _close_pools()
    sys.exit(rc)
