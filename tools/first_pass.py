"""First-pass cost of the API path in a fresh process: page-locking the model's arrays in place vs pinned staging rows.
python tools/first_pass.py cfg5 98 [0|1]   (1 = page-lock in place, the default)"""
import sys, os, time, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import workloads
from msm_we_b200 import _pinning
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.msm_we import modelWE
from msm_we_b200.stratified_clustering import StratifiedClusters

name, n_it, pin = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
_pinning.PINS.enabled = bool(pin)
cfg = dataclasses.replace(workloads.CONFIGS[name], n_iters=n_it)
means, centers = workloads.make_centers(cfg)
its = workloads.generate_host(cfg, means)
basis, target = workloads.region_bounds(cfg)
torch.zeros(1, device="cuda"); torch.cuda.synchronize()          # context creation is not what is being measured
model = modelWE()
model.initialize(workloads.to_iteration_source(its), None, "fp", basis_pcoord_bounds=basis, target_pcoord_bounds=target, tau=1.0, pcoord_ndim=1)
model.get_iterations(); model.dimReduce()
clusters = StratifiedClusters(RectilinearBinMapper(workloads.boundaries(cfg)), model, cfg.k_per_bin, [])
for b in range(cfg.n_bins):
    clusters.cluster_models[b].cluster_centers_ = centers[b].copy()
model.clusters = clusters; model.n_clusters = cfg.n_clusters
for p in range(3):
    t = [time.perf_counter()]
    model.lloyd_refine_clusters(10); torch.cuda.synchronize(); t.append(time.perf_counter())
    model.launch_ray_discretization(); torch.cuda.synchronize(); t.append(time.perf_counter())
    model.get_fluxMatrix(n_lag=0, first_iter=0); torch.cuda.synchronize(); t.append(time.perf_counter())
    print(f"pin={pin} pass {p}: lloyd {1e3*(t[1]-t[0]):.0f} ms  discretize {1e3*(t[2]-t[1]):.0f} ms  flux {1e3*(t[3]-t[2]):.0f} ms  total {1e3*(t[3]-t[0]):.0f} ms", flush=True)
