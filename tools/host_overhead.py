"""Host enqueue time vs GPU time of one hot-path step (cfg2)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads as synthetic
from msm_we_b200 import _lib, ops
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.engine import DeviceClusters
cfg = synthetic.CONFIGS["cfg2"]
dev = torch.device("cuda:0")
means, centers = synthetic.make_centers(cfg)
basis, target = synthetic.region_bounds(cfg)
eng = DeviceClusters(RectilinearBinMapper(synthetic.boundaries(cfg)), centers, {b: b for b in range(cfg.n_bins)}, basis, target, 1, device=dev)
data = synthetic.generate_device(cfg, dev, means=means)
M = cfg.n_clusters + 2
dense = torch.zeros((M, M), dtype=torch.float64, device=dev)
labels = torch.empty(2 * data["n"], dtype=torch.int64, device=dev)
def step():
    dense.zero_()
    eng.hotpath_step(data["X"], data["pcoord"], data["weights"], cfg.n_clusters, iter_offsets=data["iter_offsets"], dense=dense,
                     divisor=float(cfg.n_iters), labels_out=labels, path=_lib.ASSIGN_FP64)
for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue per step: {(t1 - t0) / 200 * 1e6:.1f} us; wall per step incl. drain: {(t2 - t0) / 200 * 1e6:.1f} us")
# CUDA graph replay of the same step
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        step()
torch.cuda.synchronize()
ref = dense.clone()
for _ in range(10): g.replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): g.replay()
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"graph replay per step: {(t1 - t0) / 200 * 1e6:.1f} us; result identical: {bool(torch.equal(ref, dense))}")
