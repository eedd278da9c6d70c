"""Host-side phase timing of one Lloyd iteration at a BASELINE shape (synchronised after every phase): where the wall
time of clustering_ops.lloyd_fit goes.  python tools/lloyd_profile.py cfg5 250"""
import sys, os, time, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads
from msm_we_b200 import _lib, clustering_ops, ops
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.engine import DeviceClusters

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 120
cfg = dataclasses.replace(workloads.CONFIGS[name], n_iters=iters)
dev = torch.device("cuda:0")
means, centers = workloads.make_centers(cfg)
basis, target = workloads.region_bounds(cfg)
eng = DeviceClusters(RectilinearBinMapper(workloads.boundaries(cfg)), centers, {b: b for b in range(cfg.n_bins)}, basis, target, 1, device=dev)
data = workloads.generate_device(cfg, dev, means=means)
N = data["n"]
X, pc = data["X"], data["pcoord"]
Xc, pc0 = X[N:], pc[:N]
bins, flags = eng.bins_and_flags(pc0)
c = eng.centers.clone()
sumK = c.shape[0]


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for it in range(12):
    if it == 6:
        c.copy_(eng.centers)           # second round from the initial centres: steady-state cost of the relocation path
    t0 = sync()
    csq = ops.centers_sqnorm(c)
    labels = ops.assign_stratified(Xc, bins, flags, c, csq, eng.bin_offset, eng.max_k, path=_lib.ASSIGN_AUTO, errors=eng.errors)
    t1 = sync()
    sum_wx, sum_w = ops.centroid_accumulate(Xc, None, labels, sumK)
    t2 = sync()
    ch = clustering_ops._relocate_empty_clusters(Xc, None, labels, c, bins, flags, eng.bin_offset, sum_wx, sum_w, None)
    t3 = sync()
    ops.lloyd_finalize(sum_wx, sum_w, c)
    t4 = sync()
    print(f"iter {it}: assign {1e3*(t1-t0):7.3f} ms  accumulate {1e3*(t2-t1):7.3f} ms  relocate {1e3*(t3-t2):7.3f} ms (changed={ch})  "
          f"finalize {1e3*(t4-t3):6.3f} ms   [{N} child frames]", flush=True)
t0 = sync()
clustering_ops.lloyd_fit(Xc, None, bins, c, eng.bin_offset, eng.max_k, 5, flags_dev=flags, errors=eng.errors)
t1 = sync()
print(f"lloyd_fit x5: {1e3*(t1-t0)/5:.3f} ms per iteration")
