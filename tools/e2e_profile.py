"""cProfile of one end-to-end pass (launch_ray_discretization + get_fluxMatrix) on host buffers."""
import cProfile, pstats, sys, time, dataclasses, io
sys.path.insert(0, ".")
import torch
import workloads as synthetic
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.msm_we import modelWE
from msm_we_b200.stratified_clustering import StratifiedClusters

cfg = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
means, centers = synthetic.make_centers(cfg)
its = synthetic.generate_host(cfg, means)
basis, target = synthetic.region_bounds(cfg)
model = modelWE()
model.initialize(synthetic.to_iteration_source(its), None, "bench", basis_pcoord_bounds=basis,
                 target_pcoord_bounds=target, tau=1.0, pcoord_ndim=1)
model.get_iterations(); model.dimReduce()
clusters = StratifiedClusters(RectilinearBinMapper(synthetic.boundaries(cfg)), model, cfg.k_per_bin, [])
for b in range(cfg.n_bins):
    clusters.cluster_models[b].cluster_centers_ = centers[b]
import os
if os.environ.get('CHUNK_MB'): clusters.cluster_args['gpu_chunk_bytes'] = int(os.environ['CHUNK_MB']) << 20
model.clusters = clusters; model.n_clusters = cfg.n_clusters

def once():
    t0 = time.perf_counter()
    model.launch_ray_discretization()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    model.get_fluxMatrix(n_lag=0, first_iter=0)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return t1 - t0, t2 - t1

once(); once()
for _ in range(3):
    print("discretize %.1f ms   flux %.1f ms" % tuple(1e3 * x for x in once()))
pr = cProfile.Profile(); pr.enable(); once(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
