"""cuBLAS DGEMM throughput on this GPU (the fp64 tensor-roofline denominator SURVEY section 8d asks for).
torch.matmul on float64 CUDA tensors dispatches to cublasDgemm; best of 10 (burst) and a 3 s back-to-back loop
(sustained), CUDA events.  Prints one JSON line."""
import json
import sys
import time

import torch


def main(n=8192):
    dev = torch.device("cuda", 0)
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = 2.0 * n ** 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); reps = 0
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            torch.matmul(a, b, out=c)
        reps += 5
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sustained = flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    print(json.dumps({"dgemm_tflops": flops / (best * 1e-3) / 1e12, "dgemm_tflops_sustained": sustained, "n": n,
                      "gpu": torch.cuda.get_device_name(0), "how": "torch.matmul float64 (cublasDgemm), CUDA events"}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8192)
