"""cProfile of one end-to-end pass (host numpy in / out) at a BASELINE shape: where the host time of the API path goes."""
import cProfile, pstats, sys, os, dataclasses, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import workloads
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.msm_we import modelWE
from msm_we_b200.stratified_clustering import StratifiedClusters

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 60
lloyd = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cfg = dataclasses.replace(workloads.CONFIGS[name], n_iters=n_it)
means, centers = workloads.make_centers(cfg)
its = workloads.generate_host(cfg, means)
basis, target = workloads.region_bounds(cfg)
model = modelWE()
model.initialize(workloads.to_iteration_source(its), None, "prof", basis_pcoord_bounds=basis, target_pcoord_bounds=target, tau=1.0, pcoord_ndim=1)
model.get_iterations(); model.dimReduce()
clusters = StratifiedClusters(RectilinearBinMapper(workloads.boundaries(cfg)), model, cfg.k_per_bin, [])
for b in range(cfg.n_bins):
    clusters.cluster_models[b].cluster_centers_ = centers[b].copy()
model.clusters = clusters; model.n_clusters = cfg.n_clusters


def once():
    t = [time.perf_counter()]
    if lloyd:
        for b in range(cfg.n_bins):
            model.clusters.cluster_models[b].cluster_centers_ = centers[b].copy()
        model.lloyd_refine_clusters(lloyd)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    model.launch_ray_discretization()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    model.get_fluxMatrix(n_lag=0, first_iter=0)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    return [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]


once(); once()
print("phases (ms) lloyd / discretize / flux:", once())
pr = cProfile.Profile(); pr.enable(); once(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
