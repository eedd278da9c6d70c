"""Kernel-level timing of the device projection (X - mean) @ components.T against numpy on the host."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from msm_we_b200 import ops

dev = torch.device("cuda:0")
for N, D_in, d_out in ((400000, 80, 13), (400000, 256, 64), (100000, 3000, 50)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    X = torch.randn((N, D_in), dtype=torch.float64, device=dev, generator=g)
    W = torch.randn((d_out, D_in), dtype=torch.float64, device=dev, generator=g)
    m = torch.randn(D_in, dtype=torch.float64, device=dev, generator=g)
    for _ in range(3): Y = ops.project(X, W, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): Y = ops.project(X, W, m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = N * (D_in + d_out) * 8 / 1e9
    n_host = min(N, 50000)
    Xh, Wh, mh = X[:n_host].cpu().numpy(), W.cpu().numpy(), m.cpu().numpy()
    t0 = time.perf_counter(); ref = (Xh - mh) @ Wh.T; th = (time.perf_counter() - t0) * N / n_host
    err = float(np.abs(Y[:n_host].cpu().numpy() - ref).max() / np.abs(ref).max())
    print(f"N={N} {D_in}->{d_out}: {ms*1e3:8.1f} us  {gb/ms*1e3:7.0f} GB/s ({gb/ms*1e3/6555.8*100:4.1f}% of HBM peak)   "
          f"numpy on the host (all cores, scaled): {th*1e3:8.1f} ms   max rel diff {err:.1e}")
