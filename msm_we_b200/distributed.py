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


_PEER_REDUCERS = {}


def peer_reducer(shape, device, group=None):
    """Cached ``PeerFluxAllreduce`` for this matrix shape (None when the ranks cannot map each other's memory)."""
    key = (tuple(shape), str(device), id(group))
    if key not in _PEER_REDUCERS:
        _PEER_REDUCERS[key] = PeerFluxAllreduce.create(shape, device, group)
    return _PEER_REDUCERS[key]


def allreduce_flux(local_sum, n_iters_total, group=None):
    """all-reduce(sum) of the per-rank un-normalised flux matrix, then the reference's ``/ nI``
    (_fluxmatrix.py:342).  CUDA tensors of ranks that can map each other's memory go through the one-kernel
    peer-memory exchange (rank-order sum); otherwise NCCL (CUDA) / gloo (CPU) all-reduce in place."""
    import torch.distributed as dist

    if local_sum.is_cuda and n_iters_total != 0 and dist.is_available() and dist.is_initialized() \
            and dist.get_world_size(group) > 1:
        red = peer_reducer(local_sum.shape, local_sum.device, group)
        if red is not None:
            red.partial.copy_(local_sum)
            # ranks gather their iterations on the host at different speeds: meet here, so the kernel's spin only has
            # to cover launch skew (its timeout means "a peer is gone", and then EVERY rank raises)
            dist.barrier(group)
            out = red.reduce(float(n_iters_total))
            red.errors.check()
            return out.clone()
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


class _RawDeviceBuffer:
    """cudaMalloc'd memory (so that its CUDA IPC handle maps exactly this buffer) viewed as a torch tensor."""

    def __init__(self, nbytes, typestr, shape):
        import ctypes as C

        from . import _lib

        p = C.c_void_p()
        _lib.check(_lib.lib.mwe_device_malloc(int(nbytes), C.byref(p)), "mwe_device_malloc")
        self.ptr = p.value
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (self.ptr, False), "version": 2}

    def free(self):
        from . import _lib

        if self.ptr:
            _lib.lib.mwe_device_free(self.ptr)
            self.ptr = None


class PeerFluxAllreduce:
    """The exchange step of the sharded flux path as ONE kernel per rank over NVLink peer memory
    (csrc/peer_reduce.cu): ``out = (partial_0 + partial_1 + ...) / divisor`` on every rank, summed in rank order
    (= iteration order, ranks hold contiguous iteration blocks), instead of NCCL all-reduce + a divide launch.

    ``partial`` is the tensor the local K3 launches accumulate into (zero it, run the step, call ``reduce``);
    ``out`` holds the result afterwards.  Needs one process per GPU on one node with peer access between all of
    them; ``PeerFluxAllreduce.create`` returns None when that is not available and the caller keeps the NCCL path.
    Handles travel once, through ``torch.distributed`` object collectives."""

    def __init__(self, shape, rank, world, device):
        """Local half only (allocate + export); nothing here talks to other ranks, so a failure cannot leave the
        ranks in mismatched collectives.  ``create`` exchanges the handles and calls ``_open_peers``."""
        import ctypes as C

        import torch

        from . import _lib, ops

        self.rank, self.world, self.device, self.shape = rank, world, device, tuple(shape)
        self.count = int(np.prod(self.shape))
        self.epoch = 0
        self._opened = []
        self._bufs = []
        self._bufs = [_RawDeviceBuffer(self.count * 8, "<f8", self.shape), _RawDeviceBuffer(self.count * 8, "<f8", self.shape),
                      _RawDeviceBuffer(2 * world * 4 + 4, "<u4", (2 * world + 1,))]
        self.partial = torch.as_tensor(self._bufs[0], device=device)
        self.out = torch.as_tensor(self._bufs[1], device=device)
        self.handles = []
        for b in self._bufs:
            h = (C.c_ubyte * 64)()
            _lib.check(_lib.lib.mwe_ipc_export(b.ptr, h), "mwe_ipc_export")
            self.handles.append(bytes(h))
        self.errors = ops.DeviceErrors(device)

    def _open_peers(self, gathered):
        import ctypes as C

        from . import _lib

        world, rank = self.world, self.rank
        ptrs = [[0] * world for _ in range(3)]
        for r in range(world):
            for k in range(3):
                if r == rank:
                    ptrs[k][r] = self._bufs[k].ptr
                else:
                    p = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(gathered[r][k])
                    _lib.check(_lib.lib.mwe_ipc_open(hb, C.byref(p)), "mwe_ipc_open")
                    self._opened.append(p.value)
                    ptrs[k][r] = p.value
        arr = C.c_void_p * world
        self._partials, self._outs, self._flagps = arr(*ptrs[0]), arr(*ptrs[1]), arr(*ptrs[2])
        self._counter = self._bufs[2].ptr + 2 * world * 4      # the spare u32 behind the flags

    @classmethod
    def create(cls, shape, device, group=None):
        """None when the ranks cannot map each other's memory (the caller keeps the NCCL path).  Every rank takes part
        in exactly two object collectives whatever fails locally: (handles or None), then (peers opened or not)."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or world > 16:
            return None
        obj = None
        try:
            obj = cls(shape, rank, world, device)
        except Exception:          # allocation / export failed on this rank
            obj = None
        gathered = [None] * world
        dist.all_gather_object(gathered, None if obj is None else obj.handles, group=group)
        ok = obj is not None and all(g is not None for g in gathered)
        if ok:
            try:
                obj._open_peers(gathered)
            except Exception:      # no peer access / IPC between these processes
                ok = False
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            if obj is not None:
                obj.close()
            return None
        return obj

    def reduce(self, divisor=0.0):
        """All ranks must call this the same number of times.  Stream-ordered; returns ``self.out``."""
        import torch

        from . import _lib

        self.epoch += 1
        _lib.check(_lib.lib.mwe_flux_peer_allreduce_f64(self._partials, self._outs, self._flagps, self.rank, self.world,
                                                        self.count, float(divisor), self.epoch, self._counter,
                                                        self.errors.counts.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream),
                   "mwe_flux_peer_allreduce_f64")
        return self.out

    def close(self):
        from . import _lib

        for p in self._opened:
            _lib.lib.mwe_ipc_close(p)
        self._opened = []
        for b in self._bufs:
            b.free()
