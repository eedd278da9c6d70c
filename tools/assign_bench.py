"""Kernel-level timing of K1 (both precision paths) on a BASELINE-shaped batch; prints GB/s and labels agreement."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dataclasses
import numpy as np, torch
import workloads as synthetic
from msm_we_b200 import _lib, ops
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.engine import DeviceClusters

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else None
cfg = synthetic.CONFIGS[name]
if iters:
    cfg = dataclasses.replace(cfg, n_iters=iters)
dev = torch.device("cuda:0")
means, centers = synthetic.make_centers(cfg)
basis, target = synthetic.region_bounds(cfg)
eng = DeviceClusters(RectilinearBinMapper(synthetic.boundaries(cfg)), centers, {b: b for b in range(cfg.n_bins)}, basis, target, 1, device=dev)
data = synthetic.generate_device(cfg, dev, means=means)
bins, flags = eng.bins_and_flags(data["pcoord"])
X = data["X"]
n2 = X.shape[0]
out = {}
for pname, path in (("fp64", _lib.ASSIGN_FP64), ("tf32x3", _lib.ASSIGN_TF32X3)):
    for _ in range(2):
        lab = ops.assign_stratified(X, bins, flags, eng.centers, eng.csq, eng.bin_offset, eng.max_k, path=path)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in evs:
        _lib.set_timing_events(a, b)
        lab = ops.assign_stratified(X, bins, flags, eng.centers, eng.csq, eng.bin_offset, eng.max_k, path=path)
        _lib.set_timing_events(None, None)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[2]
    gbs = n2 * (cfg.dim * 8 + 12) / ms / 1e6
    tf = n2 * 2.0 * cfg.k_per_bin * cfg.dim / ms / 1e9
    out[pname] = lab
    print(f"{name} {pname:7s} points={n2} D={cfg.dim} K={cfg.k_per_bin}: {ms:8.3f} ms  {gbs:7.1f} GB/s ({gbs/6555.8*100:4.1f}% HBM)  {tf:6.1f} TFLOP/s-equivalent")
print("labels identical:", bool(torch.equal(out["fp64"], out["tf32x3"])))
