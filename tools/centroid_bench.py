"""Kernel-level timing of K2 (centroid accumulation / Lloyd M step) on a cfg2-shaped batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads as synthetic
from msm_we_b200 import ops
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.engine import DeviceClusters

import dataclasses
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = synthetic.CONFIGS[name]
if len(sys.argv) > 2:
    cfg = dataclasses.replace(cfg, n_iters=int(sys.argv[2]))
dev = torch.device("cuda:0")
means, centers = synthetic.make_centers(cfg)
basis, target = synthetic.region_bounds(cfg)
eng = DeviceClusters(RectilinearBinMapper(synthetic.boundaries(cfg)), centers, {b: b for b in range(cfg.n_bins)}, basis, target, 1, device=dev)
data = synthetic.generate_device(cfg, dev, means=means)
X = data["X"]
labels, bins, flags = eng.predict(X, data["pcoord"])
sumK = eng.total
lab = labels.clone(); lab[flags != 0] = -1          # basis/target points do not enter the centroid update
w = torch.rand(X.shape[0], dtype=torch.float64, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t = timeit(lambda: ops.centroid_accumulate(X, w, lab, sumK))
n = X.shape[0]
gb = n * (cfg.dim * 8 + 8 + 8) / 1e9
print(f"{name}: centroid_accumulate N={n} D={cfg.dim} sumK={sumK}: {t*1e3:.1f} us, {gb/t*1e3:.0f} GB/s ({gb/t*1e3/6555.8*100:.1f}% of HBM peak)")
t1 = timeit(lambda: eng.predict(X, data["pcoord"]))
print(f"{name}: K0+K1 on the same batch {t1*1e3:.1f} us -> one Lloyd iteration (assign + accumulate) {(t+t1)*1e3:.1f} us")
