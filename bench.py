#!/usr/bin/env python
"""Benchmark of the discretization + flux hot path (BASELINE.json metric: WE frames/s assigned and
flux-accumulated; 1 frame = 1 WE segment in 1 iteration = 2 feature vectors assigned + 1 weighted
transition scattered).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl b200|reference]

One step = one pass of the hot path over the whole workload batch held by this rank:
K0 (bin + basis/target flags of parent and child pcoords) -> K1 (stratified assignment of parent and
child features) -> K3 (sort + segmented fp64 sum into the dense flux matrix, / nI); with N > 1 ranks
every rank owns its own iteration range (weak scaling, fixed work per GPU) and the per-rank flux
matrices are combined with one NCCL all-reduce, the path's only exchange step.

value  : frames/s with the inputs already resident in HBM (CUDA events per step, L2 flushed between
         steps, max over ranks);
e2e    : the same metric through the modelWE plugin API (launch_ray_discretization + get_fluxMatrix)
         with host numpy buffers: H2D of every feature/pcoord/weight and D2H of labels and the flux
         matrix are inside the timed region;
roofline: the dominant kernel (K1, assign_dmma_kernel), algorithmic bytes / CUDA-event time measured
         inside the timed steps on the launching stream;
cpu_baseline / --impl reference: the reference's own CPU pattern (oracle literal loop: one sklearn
         predict([x]) per segment + per-iteration scipy coo_matrix -> dense add) on the host cores.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "we_frames_per_sec_assigned_and_flux_accumulated"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-iters", type=int, default=0, help="iterations of the workload timed on the CPU (0 = auto)")
    ap.add_argument("--precision-path", default="auto", choices=["auto", "fp64", "tf32x3"],
                    help="K1 evaluation: fp64 DMMA, or tcgen05 split-TF32 candidates + fp64 re-check (same labels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's literal pattern, one process per iteration range
# ----------------------------------------------------------------------------------------------

_POOL = {}


def _close_pools():
    for pool in _POOL.values():
        pool.close()
        pool.join()
    _POOL.clear()


def cpu_reference_pass(cfg_name, n_sample_iters, n_procs):
    """Times the reference pattern on `n_sample_iters` iterations of the workload; returns
    (frames, seconds, cores).  The worker pool is created once (fork) and reused by every step."""
    import multiprocessing as mp

    chunks = np.array_split(np.arange(n_sample_iters), n_procs)
    chunks = [c for c in chunks if len(c)]
    key = (cfg_name, n_sample_iters, len(chunks))
    if key not in _POOL:
        _POOL[key] = mp.get_context("fork").Pool(len(chunks))
    pool = _POOL[key]
    t0 = time.perf_counter()
    res = pool.map(_cpu_chunk, [(cfg_name, int(c[0]), int(c[-1]) + 1) for c in chunks], chunksize=1)
    dt = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    return frames, dt, len(chunks)


_CPU_CACHE = {}


def _cpu_inputs(cfg_name, n_iters):
    """Host data for the CPU arm: the same generator and seed as the small-config tests, restricted to
    `n_iters` iterations (built once in the parent, inherited by fork)."""
    key = (cfg_name, n_iters)
    if key not in _CPU_CACHE:
        import dataclasses

        import workloads as synthetic

        cfg = dataclasses.replace(synthetic.CONFIGS[cfg_name], n_iters=n_iters)
        means, centers = synthetic.make_centers(cfg)
        its = synthetic.generate_host(cfg, means)
        _CPU_CACHE[key] = (cfg, centers, its)
    return _CPU_CACHE[key]


def _cpu_chunk(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle as O
    import workloads as synthetic

    cfg_name, lo, hi, n_total = args[0], args[1], args[2], None
    cfg, centers, its = _cpu_inputs(cfg_name, _CPU_SAMPLE[0])
    basis, target = synthetic.region_bounds(cfg)
    om = O.RectilinearBinMapperOracle(synthetic.boundaries(cfg))
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


_CPU_SAMPLE = [0]


def run_cpu_arm(cfg_name, sample_iters, cores):
    _CPU_SAMPLE[0] = sample_iters
    _cpu_inputs(cfg_name, sample_iters)  # build before forking
    from oracle import oracle as O       # import sklearn/scipy in the parent so the forked workers inherit them
    import sklearn.cluster  # noqa: F401

    O.make_fitted_minibatch(np.zeros((2, 2))).predict([[0.0, 0.0]])
    frames, dt, used = cpu_reference_pass(cfg_name, sample_iters, cores)
    return frames, dt, used


def auto_sample_iters(cfg, cores, n_steps=1, budget_s=20.0):
    """Iterations of the workload one CPU step covers: the whole (warm-up + timed) run is sized to about
    `budget_s` seconds of wall time per core set, at least one iteration per core, at most the workload."""
    per_iter = cfg.n_segs * 2 * 230e-6 + 4e-9 * (cfg.n_clusters + 2) ** 2   # measured: ~0.45 core-s per cfg2 iteration
    per_core = max(1, int(budget_s / max(n_steps, 1) / per_iter))
    return int(min(cfg.n_iters, per_core * cores))


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
                                          "--format=csv,noheader,nounits", "-lms", "10", "-f", self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            # nvidia-smi needs up to a second to start on a fresh box: wait for its first sample, else a short
            # timed region ends before anything was written
            t_end = time.time() + 5.0
            while time.time() < t_end:
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
            time.sleep(0.03)
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


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import workloads as synthetic

    cfg = synthetic.CONFIGS[args.workload]
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = args.cpu_sample_iters or auto_sample_iters(cfg, cores, args.steps + args.warmup, budget_s=120.0)
        times = []
        frames = 0
        for step in range(args.warmup + args.steps):
            frames, dt, used = run_cpu_arm(args.workload, sample, cores)
            if step >= args.warmup:
                times.append(dt)
        total = sum(times)
        val = frames * len(times) / total
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{cfg.name}: {cfg.n_iters} WE iters x {cfg.n_segs} segs x {cfg.dim}-dim, "
                                   f"{cfg.n_bins} bins x {cfg.k_per_bin} clusters/bin, fp64"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port",
                             "sample": f"{sample} of {cfg.n_iters} iterations per step ({frames} frames), literal "
                                       f"reference loop (sklearn predict([x]) per segment + scipy coo->dense add per "
                                       f"iteration), one process per iteration range (ray unavailable: multiprocessing)"},
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

    from msm_we_b200 import _lib, ops
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.engine import DeviceClusters

    means, centers = synthetic.make_centers(cfg)
    basis, target = synthetic.region_bounds(cfg)
    mapper = RectilinearBinMapper(synthetic.boundaries(cfg))
    remap = {b: b for b in range(cfg.n_bins)}
    engine = DeviceClusters(mapper, centers, remap, basis, target, 1, device=dev)
    data = synthetic.generate_device(cfg, dev, means=means, seed_offset=rank)   # this rank's iteration range
    N = data["n"]
    n_clusters = cfg.n_clusters
    M = n_clusters + 2
    X, pc, w, offs = data["X"], data["pcoord"], data["weights"], data["iter_offsets"]
    # N > 1: the per-rank matrix lives in a buffer every rank can map; one kernel per rank does the all-reduce and
    # the "/ nI" over NVLink peer memory (NCCL all-reduce + divide when the ranks cannot map each other)
    reducer = None
    if world > 1:
        from msm_we_b200.distributed import PeerFluxAllreduce
        reducer = PeerFluxAllreduce.create((M, M), dev)
    dense = reducer.partial if reducer is not None else torch.zeros((M, M), dtype=torch.float64, device=dev)
    labels = torch.empty(2 * N, dtype=torch.int64, device=dev)
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    n_iters_total = cfg.n_iters * world
    path = {"auto": _lib.ASSIGN_AUTO, "tf32x3": _lib.ASSIGN_TF32X3, "fp64": _lib.ASSIGN_FP64}[args.precision_path]
    if path == _lib.ASSIGN_AUTO:   # same rule as the library (csrc/assign.cu, resolve_assign_path)
        path = _lib.ASSIGN_TF32X3 if cfg.k_per_bin * cfg.dim >= 2048 else _lib.ASSIGN_FP64

    def step(ev=None):
        if ev is not None:
            _lib.set_timing_events(ev[0], ev[1])
        dense.zero_()
        # one C call enqueues K0 -> K1 -> K3 (-> / nI when single-GPU)
        engine.hotpath_step(X, pc, w, n_clusters, iter_offsets=offs, dense=dense,
                            divisor=float(n_iters_total) if world == 1 else 0.0, labels_out=labels, path=path)
        if ev is not None:
            _lib.set_timing_events(None, None)
        if reducer is not None:
            reducer.reduce(float(n_iters_total))
        elif world > 1:
            dist.all_reduce(dense)
            ops.divide_(dense, float(n_iters_total))

    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    for _ in range(args.warmup):
        l2_flush.zero_()
        step()
    torch.cuda.synchronize()
    engine.check_errors()

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.mark_timed(True)
    for k in range(args.steps):
        l2_flush.zero_()                       # flush L2 between timed steps (outside the events)
        evs[k][0].record()
        step(kevs[k])
        evs[k][1].record()
    torch.cuda.synchronize()
    clocks.mark_timed(False)
    clocks.__exit__()
    if world > 1:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kevs) / args.steps
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    engine.check_errors()
    if reducer is not None:
        reducer.errors.check()
    frames_per_step = N * world
    value = frames_per_step * args.steps / (total_ms * 1e-3)

    # roofline of the dominant kernel (K1): algorithmic bytes per launch = per point D*8 (features) +
    # 4 (bucket index) + 8 (int64 label); 2N points per launch
    peak, peak_src = load_peaks()
    alg_bytes = 2 * N * (cfg.dim * 8 + 4 + 8)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "assign_tc_kernel" if path == _lib.ASSIGN_TF32X3 else
                ("assign_dmma_resident_kernel" if (cfg.k_per_bin <= 64 and cfg.dim % 2 == 0) else "assign_dmma_kernel"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "fp64_tflops": 2 * N * 2.0 * cfg.k_per_bin * cfg.dim / (kernel_ms * 1e-3) / 1e12}

    # DRAM bytes per launch of that kernel from the committed ncu capture of the same workload (None if none)
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "k1_dram_traffic.json")) as f:
            entry = json.load(f).get(f"{cfg.name}/{roofline['kernel']}")
        if entry:
            roofline["traffic"] = entry["bytes"]
            roofline["traffic_source"] = entry["source"]
    except (OSError, ValueError):
        pass

    # ---------------------------------------------------------------- e2e through the plugin API
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(cfg, rank, world, dev, max(2, min(args.steps, 5)))

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample_iters or auto_sample_iters(cfg, cores)
        frames, dt, used = run_cpu_arm(args.workload, sample, cores)
        cpu = {"value": frames / dt, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"{sample} of {cfg.n_iters} iterations ({frames} frames) of the same workload shape, literal "
                         f"reference loop, {dt:.1f} s wall"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{cfg.name}: {cfg.n_iters} WE iters x {cfg.n_segs} segs x {cfg.dim}-dim, "
                                   f"{cfg.n_bins} bins x {cfg.k_per_bin} clusters/bin, fp64 (per GPU; iteration-range "
                                   f"sharded, flux all-reduced)",
                       "frames_per_step": frames_per_step, "l2": "flushed between timed steps (256 MiB memset)",
                       "exchange": ("none (1 GPU)" if world == 1 else "one peer-memory kernel per rank (rank-order sum + / nI over NVLink)"
                                    if reducer is not None else "NCCL all-reduce + divide"),
                       "precision_path": ("tcgen05 split-TF32 candidate pass + fp64 re-check of near-ties (labels identical "
                                          "to the fp64 path)") if path == _lib.ASSIGN_TF32X3 else "fp64 DMMA"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks.summary(),
            "gpu_launches": LAUNCHES_PER_STEP_STATIC(cfg, path == _lib.ASSIGN_TF32X3) * args.steps,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        if reducer is not None:
            reducer.close()
        dist.destroy_process_group()
    return 0


def LAUNCHES_PER_STEP_STATIC(cfg, tc_path=False):
    """Kernels of ours per step, counted from the launch sequences in csrc/ (and confirmed by the ncu launch list
    in profiles/): K0 bin_flags (1) + K1 scan, scatter, main kernel, re-check (4; +3 centre-preparation kernels on
    the tcgen05 path) + K3 keys (1), radix sort = 1 histogram + one scatter per 8-bit pass, mark+scan (1), group sum (1), cell sum (1) + the final divide or,
    with N > 1, the peer-memory exchange kernel (1)."""
    M = cfg.n_clusters + 2
    bits = int(np.ceil(np.log2(M * M + 1)))
    passes = (bits + 7) // 8
    scans = 0          # the scatter kernels scan the per-CTA histograms themselves for every grid the sort launches
    return 1 + 4 + (3 if tc_path else 0) + 1 + (1 + passes + scans) + 3 + 1


def run_e2e(cfg, rank, world, dev, steps):
    """frames/s through the public modelWE API with HOST buffers (pinned staging inside the API)."""
    import torch
    import torch.distributed as dist

    import workloads as synthetic
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.msm_we import modelWE
    from msm_we_b200.stratified_clustering import StratifiedClusters
    import dataclasses

    # host copy of a bounded number of iterations of the same shape (host generation is slow for big configs)
    n_it = min(cfg.n_iters, 200)
    hcfg = dataclasses.replace(cfg, n_iters=n_it, seed=cfg.seed + rank)
    means, centers = synthetic.make_centers(cfg)
    its = synthetic.generate_host(hcfg, means)
    basis, target = synthetic.region_bounds(cfg)
    model = modelWE()
    model.initialize(synthetic.to_iteration_source(its), None, "bench", basis_pcoord_bounds=basis,
                     target_pcoord_bounds=target, tau=1.0, pcoord_ndim=1)
    model.get_iterations()
    model.dimReduce()
    clusters = StratifiedClusters(RectilinearBinMapper(synthetic.boundaries(cfg)), model, cfg.k_per_bin, [])
    for b in range(cfg.n_bins):
        clusters.cluster_models[b].cluster_centers_ = centers[b]
    model.clusters = clusters
    model.n_clusters = cfg.n_clusters
    model.pre_discretization_model = model   # skip the deepcopy of the whole in-memory data set

    def once():
        model.launch_ray_discretization()
        model.get_fluxMatrix(n_lag=0, first_iter=0)
        return model.fluxMatrixRaw

    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    frames = (n_it - 1) * cfg.n_segs          # launch_ray_discretization covers range(1, maxIter)
    M = cfg.n_clusters + 2
    h2d = frames * (2 * (cfg.dim + 1) * 8) + frames * (2 * 8 + 8 + 2 * 8)
    d2h = frames * 2 * (8 + 4 + 1) + M * M * 8
    return {"value": frames * world * steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "frames_per_step": frames * world,
            "api": "modelWE.launch_ray_discretization + get_fluxMatrix, numpy in / numpy out"}


if __name__ == "__main__":
    try:
        rc = main()
    finally:
        _close_pools()
    sys.exit(rc)
