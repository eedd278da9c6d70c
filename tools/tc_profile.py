"""Per-role wait-cycle breakdown of the tcgen05 assignment kernel on a cfg2-shaped batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import workloads as synthetic
from msm_we_b200 import _lib, ops
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.engine import DeviceClusters

cfg = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda:0")
means, centers = synthetic.make_centers(cfg)
basis, target = synthetic.region_bounds(cfg)
eng = DeviceClusters(RectilinearBinMapper(synthetic.boundaries(cfg)), centers, {b: b for b in range(cfg.n_bins)}, basis, target, 1, device=dev)
data = synthetic.generate_device(cfg, dev, means=means, iters=min(cfg.n_iters, 200))
bins, flags = eng.bins_and_flags(data["pcoord"])
prof = torch.zeros(20, dtype=torch.int64, device=dev)
for rep in range(3):
    prof.zero_()
    _lib.lib.mwe_debug_set_tc_profile(prof.data_ptr())
    ops.assign_stratified(data["X"], bins, flags, eng.centers, eng.csq, eng.bin_offset, eng.max_k, path=_lib.ASSIGN_TF32X3)
    torch.cuda.synchronize()
    _lib.lib.mwe_debug_set_tc_profile(None)
v = prof.cpu().numpy().reshape(5, 4) / 148.0
for name, row in zip(["mma", "centre", "epilogue", "raw", "convert"], v):
    print(f"{name:9s} total {row[0]:9.0f} cyc  wait0 {row[1]:9.0f}  wait1 {row[2]:9.0f}  wait2 {row[3]:9.0f}")
