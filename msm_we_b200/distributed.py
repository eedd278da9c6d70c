"""Multi-GPU plumbing of the hot path: one process per GPU, WE iterations sharded across ranks.

reference: the path shards by WE iteration -- Ray runs one task per iteration for discretization
(msm_we/_hamsm/_clustering.py:1183-1192) and for the flux matrix (_fluxmatrix.py:275-300) and adds the
per-iteration matrices on the driver (:311-327).  Here every rank owns a contiguous block of iterations
(balanced by segment count, since the number of segments grows during a WE run), labels never leave the
rank, and the only exchange step is an all-reduce(sum) of the un-normalised flux matrix (NCCL over
NVLink on GPUs; gloo in the CPU tests), followed by ``/ nI`` with the GLOBAL iteration count.
"""
from __future__ import annotations

import numpy as np


def partition_iterations(iters, seg_counts, world_size):
    """Split ``iters`` (in order) into ``world_size`` contiguous blocks with near-equal segment totals.
    ``seg_counts[i]`` is the number of segments of ``iters[i]``.  Blocks may be empty when there are fewer
    iterations than ranks."""
    iters = list(iters)
    counts = np.asarray(seg_counts, dtype=np.float64)
    if len(iters) != len(counts):
        raise ValueError("iters and seg_counts must have the same length")
    total = counts.sum()
    bounds = [0]
    cum = np.cumsum(counts)
    for r in range(1, world_size):
        target = total * r / world_size
        # first index whose cumulative count reaches the target, never moving backwards
        idx = int(np.searchsorted(cum, target, side="left")) + 1 if total > 0 else 0
        idx = min(max(idx, bounds[-1]), len(iters))
        bounds.append(idx)
    bounds.append(len(iters))
    return [iters[bounds[r]:bounds[r + 1]] for r in range(world_size)]


def allreduce_flux(local_sum, n_iters_total, group=None):
    """all-reduce(sum) of the per-rank un-normalised flux matrix, then the reference's ``/ nI``
    (_fluxmatrix.py:342).  ``local_sum`` is a torch tensor (CUDA -> NCCL, CPU -> gloo), reduced in place."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_sum, op=dist.ReduceOp.SUM, group=group)
    if n_iters_total == 0:
        return local_sum / 0.0
    if local_sum.is_cuda:
        from . import ops

        return ops.divide_(local_sum, float(n_iters_total))
    return local_sum.div_(float(n_iters_total))


def get_fluxMatrix_sharded(model, n_lag=0, first_iter=1, last_iter=None, iters_to_use=None, group=None,
                           local_flux_fn=None):
    """``modelWE.get_fluxMatrix`` across the ranks of ``group``: every rank scatters its own iteration block
    (K0 + K3 on its GPU), the matrices are all-reduced, every rank ends with the same ``fluxMatrixRaw``.
    ``local_flux_fn(model, iters) -> torch tensor`` is injectable for the CPU (gloo) tests."""
    import torch.distributed as dist

    world, rank = 1, 0
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    model._fluxMatrixParams = [n_lag, first_iter, last_iter, iters_to_use]
    if iters_to_use is None:
        if last_iter is None:
            last_iter = model.maxIter
        iters_to_use = range(first_iter + 1, last_iter)
    iters_to_use = list(iters_to_use)
    model.n_lag = n_lag
    model.errorWeight = 0.0
    model.errorCount = 0
    seg_counts = [model.numSegments[i - 1] if 0 < i <= len(model.numSegments) else 0 for i in iters_to_use]
    mine = partition_iterations(iters_to_use, seg_counts, world)[rank]
    fn = local_flux_fn if local_flux_fn is not None else (lambda m, its: m._flux_device(its))
    local = fn(model, mine)
    out = allreduce_flux(local, len(iters_to_use), group)
    model.fluxMatrixRaw = out.cpu().numpy()
    return model.fluxMatrixRaw
