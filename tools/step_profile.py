"""Where one bench step (10 Lloyd iterations + assign + flux) of a config-5 SHARD spends its wall time on one GPU:
un-synchronised step time, then the same step with a synchronise after every phase.
python tools/step_profile.py cfg5 500"""
import sys, os, time, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads
from msm_we_b200 import _lib, clustering_ops, ops
from msm_we_b200.binning import RectilinearBinMapper
from msm_we_b200.engine import DeviceClusters

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 500
cfg = dataclasses.replace(workloads.CONFIGS[name], n_iters=iters)
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lrank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrank)
dev = torch.device(f"cuda:{lrank}")
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    flux_group = dist.new_group()
side = torch.cuda.Stream(device=dev)
means, centers = workloads.make_centers(cfg)
basis, target = workloads.region_bounds(cfg)
eng = DeviceClusters(RectilinearBinMapper(workloads.boundaries(cfg)), centers, {b: b for b in range(cfg.n_bins)}, basis, target, 1, device=dev)
data = workloads.generate_device(cfg, dev, means=means, seed_offset=int(os.environ.get("SEED_OFFSET", rank)))
N = data["n"]
X, pc, w, offs = data["X"], data["pcoord"], data["weights"], data["iter_offsets"]
Xc, pc0 = X[N:], pc[:N]
c0 = eng.centers.clone()
M = cfg.n_clusters + 2
dense = torch.zeros((M, M), dtype=torch.float64, device=dev)
labels_out = torch.empty(2 * N, dtype=torch.int64, device=dev)
path = _lib.ASSIGN_TF32X3


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def step(phases=None):
    def ph(tag):
        if phases is not None:
            phases.append((tag, sync()))
    ph("start")
    dense.zero_()
    eng.centers.copy_(c0)
    bins_p, flags_p = eng.bins_and_flags(pc0)
    ph("prep")
    clustering_ops.lloyd_fit(Xc, None, bins_p, eng.centers, eng.bin_offset, eng.max_k, 10, group=group, flags_dev=flags_p, path=path, errors=eng.errors)
    ph("lloyd x10")
    eng.csq = ops.centers_sqnorm(eng.centers)
    eng.hotpath_step(X, pc, w, cfg.n_clusters, iter_offsets=offs, dense=dense, divisor=float(iters) if world == 1 else 0.0, labels_out=labels_out, path=path)
    ph("assign+flux")
    if world > 1:
        dist.all_reduce(dense, group=flux_group)
        ops.divide_(dense, float(iters * world))
        ph("flux exchange")


for _ in range(3):
    step()
t0 = sync()
for _ in range(5):
    step()
t1 = sync()
if rank == 0:
    print(f"{name} world {world}, {iters} its per rank ({N} frames): step {1e3 * (t1 - t0) / 5:.3f} ms un-synchronised", flush=True)
ph = []
step(ph)
if rank == 0:
    print("phases: " + "  ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}ms" for a, b in zip(ph[:-1], ph[1:])), flush=True)
if rank == 0:
    os.environ["MWE_RELOC_DEBUG"] = "1"
t0 = sync()
step()
t1 = sync()
if rank == 0:
    print(f"step with relocation marks: {1e3 * (t1 - t0):.3f} ms")
# one Lloyd round with a synchronise after every phase, on every rank
eng.centers.copy_(c0)
bins_p, flags_p = eng.bins_and_flags(pc0)
sumK = eng.centers.shape[0]
sums = torch.empty(sumK * (cfg.dim + 1), dtype=torch.float64, device=dev)
os.environ.pop("MWE_RELOC_DEBUG", None)
acc = [0.0] * 5
per_it = []
for it in range(10):
    t = [sync()]
    labels = ops.assign_stratified(Xc, bins_p, flags_p, eng.centers, ops.centers_sqnorm(eng.centers), eng.bin_offset, eng.max_k, path=path, errors=eng.errors)
    t.append(sync())
    sum_wx, sum_w = ops.centroid_accumulate(Xc, None, labels, sumK, out=sums)
    t.append(sync())
    if world > 1:
        dist.all_reduce(sums, group=group)
    t.append(sync())
    clustering_ops._relocate_empty_clusters(Xc, None, labels, eng.centers, bins_p, flags_p, eng.bin_offset, sum_wx, sum_w, group)
    t.append(sync())
    ops.lloyd_finalize(sum_wx, sum_w, eng.centers)
    t.append(sync())
    for k in range(5):
        acc[k] += t[k + 1] - t[k]
    per_it.append(f"{1e3 * (t[2] - t[1]):.2f}")
print(f"rank {rank}: 10 synchronised Lloyd iterations: assign {1e3*acc[0]:.2f}  accumulate {1e3*acc[1]:.2f}  all-reduce {1e3*acc[2]:.2f}  relocate {1e3*acc[3]:.2f}  finalize {1e3*acc[4]:.2f}  total {1e3*sum(acc):.2f} ms; accumulate per iteration: {' '.join(per_it)}", flush=True)
if world > 1:
    # the Lloyd exchange alone: 10 all-reduces of the partial-sum buffer, back to back
    buf = torch.zeros(eng.centers.shape[0] * (cfg.dim + 1), dtype=torch.float64, device=dev)
    for _ in range(3):
        dist.all_reduce(buf, group=group)
    t0 = sync()
    for _ in range(10):
        dist.all_reduce(buf, group=group)
    t1 = sync()
    dist.all_reduce(dense, group=flux_group)
    t2 = sync()
    dist.all_reduce(dense, group=flux_group)
    t3 = sync()
    if rank == 0:
        print(f"all-reduce of the {buf.numel() * 8 / 1e6:.1f} MB partial sums: {1e3 * (t1 - t0) / 10:.3f} ms each; of the {dense.numel() * 8 / 1e6:.0f} MB flux matrix: {1e3 * (t3 - t2):.3f} ms")
    dist.destroy_process_group()
