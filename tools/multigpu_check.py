"""Run under torchrun with >= 2 ranks (one GPU each): the sharded hot path against the single-GPU result.

  * Lloyd: every rank holds an iteration range of the child frames; ``lloyd_fit(group=...)`` all-reduces the partial
    sums (NCCL) -- centroids must equal the single-GPU fit to 1e-12, labels exactly;
  * flux: ``get_fluxMatrix_sharded`` (K0 + K3 per rank, one exchange step) must equal ``modelWE.get_fluxMatrix`` on one
    GPU to 1e-12 with the same sparsity pattern.
Prints one line per check; exit code 0 only if all hold.  Used by tests/test_multigpu.py (skipped on 1-GPU boxes).
"""
import dataclasses
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    import workloads
    from msm_we_b200 import clustering_ops, ops
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.distributed import get_fluxMatrix_sharded
    from msm_we_b200.engine import DeviceClusters
    from msm_we_b200.msm_we import modelWE
    from msm_we_b200.stratified_clustering import StratifiedClusters

    cfg = dataclasses.replace(workloads.CONFIGS["cfg2"], n_iters=24, n_segs=600, n_bins=10)
    means, centers = workloads.make_centers(cfg)
    its = workloads.generate_host(cfg, means)            # identical on every rank (seeded)
    basis, target = workloads.region_bounds(cfg)
    mapper = RectilinearBinMapper(workloads.boundaries(cfg))
    ok = True

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    # ---- Lloyd, sharded by iteration range -------------------------------------------------------
    def lloyd(iter_lo, iter_hi, group):
        eng = DeviceClusters(mapper, centers, {b: b for b in range(cfg.n_bins)}, basis, target, 1, device=dev)
        X = t(np.concatenate([its[i]["child"] for i in range(iter_lo, iter_hi)]))
        pc0 = t(np.concatenate([its[i]["pcoord0"] for i in range(iter_lo, iter_hi)]))
        bins, flags = eng.bins_and_flags(pc0)
        c = eng.centers.clone()
        clustering_ops.lloyd_fit(X, None, bins, c, eng.bin_offset, eng.max_k, 6, group=group, flags_dev=flags, errors=eng.errors)
        labels = ops.assign_stratified(X, bins, flags, c, ops.centers_sqnorm(c), eng.bin_offset, eng.max_k)
        eng.check_errors()
        return c.cpu().numpy(), labels.cpu().numpy()

    n = cfg.n_iters
    lo, hi = n * rank // world, n * (rank + 1) // world
    c_sh, l_sh = lloyd(lo, hi, dist.group.WORLD)
    c_one, l_one = lloyd(0, n, None)                      # every rank repeats the single-GPU fit on its own GPU
    rel = np.abs(c_sh - c_one).max() / np.abs(c_one).max()
    off = sum(len(its[i]["child"]) for i in range(lo))
    same_labels = np.array_equal(l_sh, l_one[off:off + len(l_sh)])
    good = rel < 1e-12 and same_labels
    ok &= good
    print(f"[rank {rank}] sharded Lloyd vs single GPU: max rel centroid diff {rel:.2e}, labels identical: {same_labels}", flush=True)

    # ---- flux through the model API --------------------------------------------------------------
    model = modelWE()
    model.initialize(workloads.to_iteration_source(its), None, "mg", basis_pcoord_bounds=basis, target_pcoord_bounds=target,
                     tau=1.0, pcoord_ndim=1)
    model.get_iterations()
    model.dimReduce()
    clusters = StratifiedClusters(mapper, model, cfg.k_per_bin, [])
    for b in range(cfg.n_bins):
        clusters.cluster_models[b].cluster_centers_ = centers[b]
    model.clusters = clusters
    model.n_clusters = cfg.n_clusters
    model.launch_ray_discretization()
    model.get_fluxMatrix(0)
    one = model.fluxMatrixRaw.copy()
    sharded = get_fluxMatrix_sharded(model, 0)
    same_pattern = np.array_equal(one == 0, sharded == 0)
    relf = np.abs(one - sharded)[one != 0].max() / np.abs(one[one != 0]).min() if (one != 0).any() else 0.0
    close = np.allclose(one, sharded, rtol=1e-12, atol=0)
    ok &= bool(same_pattern and close)
    print(f"[rank {rank}] get_fluxMatrix_sharded vs single GPU: same pattern {same_pattern}, allclose(1e-12) {close} "
          f"(worst abs diff / smallest entry {relf:.1e})", flush=True)

    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"ALL CHECKS PASSED: {bool(flag.item())}", flush=True)
    dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
