"""Per-phase clock cycles of the resident-centre fp64 K1 kernel (mwe_debug_set_k1_profile), cfg2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads as synthetic
from msm_we_b200 import _lib, ops
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.engine import DeviceClusters

cfg = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda:0")
means, centers = synthetic.make_centers(cfg)
basis, target = synthetic.region_bounds(cfg)
eng = DeviceClusters(RectilinearBinMapper(synthetic.boundaries(cfg)), centers, {b: b for b in range(cfg.n_bins)}, basis, target, 1, device=dev)
data = synthetic.generate_device(cfg, dev, means=means)
bins, flags = eng.bins_and_flags(data["pcoord"])
X = data["X"]
run = lambda: ops.assign_stratified(X, bins, flags, eng.centers, eng.csq, eng.bin_offset, eng.max_k, path=_lib.ASSIGN_FP64)
for _ in range(3): run()
prof = torch.zeros(8, dtype=torch.int64, device=dev)
_lib.lib.mwe_debug_set_k1_profile(prof.data_ptr())
run(); torch.cuda.synchronize()
_lib.lib.mwe_debug_set_k1_profile(None)
p = prof.cpu().tolist()
groups = max(p[4], 1)
sms = torch.cuda.get_device_properties(0).multi_processor_count
print(f"groups {p[4]} ({p[4]/sms:.0f} per SM)")
for name, v in zip(("consumer: ticket -> data ready", "consumer: metadata + accumulator seed", "consumer: k loop (LDS + DMMA)", "consumer: fold + epilogue"), p[:4]):
    print(f"  {name:40s} {v/groups:8.0f} cycles per group ({v/sum(p[:4])*100:4.1f}% of consumer time)")
for name, v in zip(("producer: waiting for a free buffer", "producer: claim + copy issue"), p[5:7]):
    print(f"  {name:40s} {v/groups:8.0f} cycles per group")
