"""torchrun --nproc-per-node N tools/peer_reduce_test.py : peer-memory all-reduce vs NCCL (values + time)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from msm_we_b200 import ops
from msm_we_b200.distributed import PeerFluxAllreduce

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
M = 602
red = PeerFluxAllreduce.create((M, M), dev)
if rank == 0:
    print("peer path available:", red is not None, flush=True)
if red is None:
    dist.destroy_process_group(); sys.exit(0)
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
ok = True
for it in range(5):
    x = torch.rand((M, M), dtype=torch.float64, device=dev, generator=g)
    red.partial.copy_(x)
    out = red.reduce(7.0).clone()
    ref = x.clone(); dist.all_reduce(ref); ref /= 7.0
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x)
    acc = parts[0].cpu().numpy()
    for p in parts[1:]:
        acc = acc + p.cpu().numpy()
    acc = acc / 7.0                                          # rank-order sum and true division, as the serial reference
    ok &= bool(np.array_equal(out.cpu().numpy(), acc))
    ok &= bool(torch.allclose(out, ref, rtol=1e-14, atol=1e-300))
red.errors.check()
def timeit(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
y = torch.rand((M, M), dtype=torch.float64, device=dev)
def nccl():
    dist.all_reduce(y); ops.divide_(y, 1.0000001)
t_peer = timeit(lambda: red.reduce(7.0))
t_nccl = timeit(nccl)
res = torch.tensor([float(ok), t_peer, t_nccl], device=dev, dtype=torch.float64)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: values identical on all ranks: {bool(res[0].item())}; peer kernel {t_peer:.1f} us, NCCL all-reduce + divide {t_nccl:.1f} us", flush=True)
dist.barrier()
red.close()
dist.destroy_process_group()
