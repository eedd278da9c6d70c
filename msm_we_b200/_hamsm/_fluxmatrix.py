"""Flux-matrix mixin: the reference's method names and results, K3 arithmetic.

reference: msm_we/_hamsm/_fluxmatrix.py -- ``get_iter_fluxMatrix`` (:21-72), ``build_flux_matrix``
(:97-164), ``build_flux_matrix_remote`` (:74-95), ``get_fluxMatrix`` (:166-345).

``get_fluxMatrix`` gathers (parent label, child label, parent pcoord, child pcoord, weight) of every
requested iteration into pinned buffers and runs K0 (basis/target flags) + K3 (sort + segmented fp64
sum) over whole chunks of iterations; the per-iteration dense ``(n+2)^2`` matrix of the reference
(65 ms per iteration at n=2500) never exists.  Sums are formed in the reference's serial order
(within an iteration in segment order, then iteration by iteration), so one GPU reproduces the serial
path (``use_ray=False``) bit for bit; ``use_ray`` is accepted and ignored.
"""
from __future__ import annotations

import numpy as np

from .._logging import log, ProgressBar

DEFAULT_FLUX_CHUNK = 1 << 26  # transitions per launch sequence
_FLUX_STAGE = {"host": None, "pending": None}   # pinned staging rows, reused across calls
_RESULT_POOL = []                               # page-locked result matrices: (torch tensor, numpy view)


def _to_host(dense):
    """The dense flux matrix as a numpy array.  Small matrices: a plain copy.  Large ones (config 5: 10,002^2 fp64 =
    800 MB) go into PAGE-LOCKED host memory, which the device fills at link speed (a pageable copy of that size runs at
    ~2 GB/s and was 75 % of an end-to-end pass); the array handed out is a view of that buffer, and a buffer is reused
    only once nothing outside this pool references its previous view."""
    import sys

    import torch

    if dense.numel() * 8 < (64 << 20):
        return dense.cpu().numpy()
    shape = tuple(dense.shape)
    slot = None
    for t, view in _RESULT_POOL:
        # references: the pool's tuple + this loop variable + getrefcount's argument
        if tuple(t.shape) == shape and sys.getrefcount(view) <= 3:
            slot = (t, view)
            break
    if slot is None:
        t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        slot = (t, t.numpy())
        _RESULT_POOL.append(slot)
        del _RESULT_POOL[:-4]            # at most four buffers stay page-locked
    slot[0].copy_(dense, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return slot[1]


class _RemoteShim:
    def __init__(self, fn):
        self._fn = fn

    def __get__(self, obj, objtype=None):
        return self

    def __call__(self, *a, **k):
        return self._fn(*a, **k)

    def remote(self, *a, **k):
        return self._fn(*a, **k)


def _flags_from_indices(n, ind_start_in_basis, ind_end_in_basis, ind_end_in_target):
    f0 = np.zeros(n, dtype=np.uint8)
    f1 = np.zeros(n, dtype=np.uint8)
    f0[ind_start_in_basis] |= 1
    f1[ind_end_in_basis] |= 1
    f1[ind_end_in_target] |= 2
    return f0, f1


def _build_flux_matrix(n_clusters, index_pairs, ind_start_in_basis, ind_end_in_basis, ind_end_in_target,
                       transition_weights):
    """reference: _fluxmatrix.py:97-164.  Returns a ``scipy.sparse.coo_matrix`` of shape
    ``(n_clusters+2, n_clusters+2)`` (duplicates already summed, entries in row-major order); scipy is
    only the container the reference's callers expect (``.todense()``), the scatter ran on the GPU."""
    import torch
    from scipy.sparse import coo_matrix

    from .. import ops
    from ..engine import require_cuda

    dev = require_cuda()
    try:
        start_cluster, end_cluster = np.asarray(index_pairs).T.copy()
    except Exception as e:
        log.error(index_pairs)
        raise e
    n = start_cluster.shape[0]
    M = n_clusters + 2
    if n == 0:
        return coo_matrix((M, M), dtype=np.float64)
    f0, f1 = _flags_from_indices(n, ind_start_in_basis, ind_end_in_basis, ind_end_in_target)
    w = np.ascontiguousarray(transition_weights, dtype=np.float64)
    if w.shape[0] != n:
        raise ValueError("row, column, and data array must all be the same length")
    errors = ops.DeviceErrors(dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    _, (r, c, v, nnz) = ops.flux_accumulate(t(start_cluster.astype(np.int64)), t(end_cluster.astype(np.int64)), t(w),
                                            int(n_clusters), flag0=t(f0), flag1=t(f1), want_coo=True, errors=errors)
    k = int(nnz.item())
    try:
        errors.check()
    except ValueError as e:
        log.error(f"Iter_fluxmatrix failed. Transition was from {start_cluster} -> {end_cluster} "
                  f"\n\t(Total {n_clusters + 2} clusters)\n\t(End in target: {ind_end_in_target})"
                  f"\n\t(Weights len: {len(transition_weights)})")
        raise e
    return coo_matrix((v[:k].cpu().numpy(), (r[:k].cpu().numpy(), c[:k].cpu().numpy())), shape=(M, M))


def _build_flux_matrix_remote(n_clusters, index_pairs, ind_start_in_basis, ind_end_in_basis, ind_end_in_target,
                              transition_weights, n_iter):
    """reference: _fluxmatrix.py:74-95."""
    return (_build_flux_matrix(n_clusters, index_pairs, ind_start_in_basis, ind_end_in_basis, ind_end_in_target,
                               transition_weights), n_iter)


class FluxMatrixMixin:
    fluxMatrixRaw = None
    fluxMatrix = None

    build_flux_matrix = staticmethod(_build_flux_matrix)
    build_flux_matrix_remote = _RemoteShim(_build_flux_matrix_remote)

    def _gather_flux_inputs(self, n_iter):
        """Host-side per-iteration inputs, exactly what the reference collects (:23-30, :277-288)."""
        self.load_iter_data(n_iter)
        parent_pcoords = self.pcoord0List
        child_pcoords = self.pcoord1List
        if self.nSeg == 0:
            self.get_transition_data_lag0()
            transition_weights = self.transitionWeights.copy()
        else:
            # same values as get_transition_data_lag0().transitionWeights (NaN-coordinate segments -> 0),
            # without rebuilding the coordinate-pair array the reference reloads only for this purpose
            transition_weights = self.iter_transition_weights(n_iter)
            self.transitionWeights = transition_weights
        index_pairs = np.asarray(self.pair_dtrajs[n_iter - 1])
        return index_pairs, parent_pcoords, child_pcoords, transition_weights

    def _gather_flux_inputs_lean(self, n_iter):
        """Same four arrays as ``_gather_flux_inputs`` for an iteration that is not the last one of a pass: the
        per-iteration model attributes (``n_iter``, ``weightList``, ``pcoord0List`` ...) are overwritten by the
        next iteration anyway, so only ``seg_weights[n_iter]`` (which persists in the reference, _data.py:915)
        is recorded.  Falls back to the full version for missing / empty iterations."""
        src = self.iteration_source
        if not src.has(n_iter):
            return self._gather_flux_inputs(n_iter)
        rec = src.get(n_iter)
        if rec.weights.shape[0] == 0:
            return self._gather_flux_inputs(n_iter)
        self.seg_weights[n_iter] = rec.weights.copy()
        w = rec.weights
        bad = self.iter_nan_segments(n_iter)
        if bad.shape[0] > 0:
            w = w.copy()
            w[bad] = 0.0
        P = self.pcoord_ndim
        return np.asarray(self.pair_dtrajs[n_iter - 1]), rec.pcoord0[:, :P], rec.pcoord1[:, :P], w

    def _flux_device(self, iters, progress=None, task=None):
        """Dense un-normalised sum over ``iters`` on the device, serial-order association."""
        import torch

        from .. import ops
        from ..engine import require_cuda

        dev = require_cuda()
        M = self.n_clusters + 2
        P = self.pcoord_ndim
        dense = torch.zeros((M, M), dtype=torch.float64, device=dev)
        errors = ops.DeviceErrors(dev)
        mapper = ops.MapperSpec.precomputed(1)
        chunk = int(getattr(self, "flux_chunk_transitions", DEFAULT_FLUX_CHUNK))
        # Staged as five contiguous field arrays (parent pcoord | child pcoord | weight | parent label | child label):
        # every per-iteration write is one contiguous copy, and the device needs no slicing / dtype conversion kernels.
        stagebuf = _FLUX_STAGE          # module-level: a model stays picklable / deep-copyable
        lens = []
        n = 0
        fields = (("p0", (P,), torch.float64), ("p1", (P,), torch.float64), ("w", (), torch.float64),
                  ("ls", (), torch.int64), ("le", (), torch.int64))

        def views(need):
            """numpy views of the pinned field arrays, grown (contents kept) to hold ``need`` transitions."""
            if stagebuf["pending"] is not None:
                stagebuf["pending"].synchronize()    # the H2D copies that last read the buffers have finished
                stagebuf["pending"] = None
            host = stagebuf["host"]
            if not isinstance(host, dict) or host["P"] != P or host["cap"] < need:
                old = host if isinstance(host, dict) and host["P"] == P else None
                cap = max(need, 1 << 14, 2 * (old["cap"] if old else 0))
                host = {"P": P, "cap": cap}
                for name, shape, dtype in fields:
                    host[name] = torch.empty((cap,) + shape, dtype=dtype, pin_memory=True)
                    if old is not None and n > 0:
                        host[name][:n] = old[name][:n]
                host["np"] = {name: host[name].numpy() for name, _, _ in fields}
                stagebuf["host"] = host
            return host["np"]

        def flush():
            nonlocal n
            if not lens:
                return
            if n > 0:
                host = stagebuf["host"]
                d = {name: host[name][:n].to(dev, non_blocking=True) for name, _, _ in fields}
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                stagebuf["pending"] = ev
                offs = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
                zeros = torch.zeros(n, dtype=torch.int32, device=dev)
                _, f0 = ops.bin_flags(d["p0"], mapper, self.basis_pcoord_bounds, self.target_pcoord_bounds, errors=errors,
                                      bin_out=zeros)
                _, f1 = ops.bin_flags(d["p1"], mapper, self.basis_pcoord_bounds, self.target_pcoord_bounds, errors=errors,
                                      bin_out=zeros)
                ops.flux_accumulate(d["ls"], d["le"], d["w"], int(self.n_clusters), flag0=f0, flag1=f1, iter_offsets=offs,
                                    dense=dense, errors=errors)
            lens.clear()
            n = 0

        iters = list(iters)
        for k, iS in enumerate(iters):
            # the last iteration goes through the reference's full loader, so the model is left in the state the
            # reference leaves it in
            index_pairs, p0, p1, w = (self._gather_flux_inputs if k == len(iters) - 1 else self._gather_flux_inputs_lean)(iS)
            s = w.shape[0]
            if s > 0:
                index_pairs = np.asarray(index_pairs)
                if index_pairs.shape[0] != s:
                    raise ValueError("row, column, and data array must all be the same length")
                if index_pairs.ndim != 2 or index_pairs.shape[1] != 2:
                    raise ValueError("pair_dtrajs entries must be [S, 2]")
                h = views(n + s)
                h["p0"][n:n + s] = p0.reshape(s, P)
                h["p1"][n:n + s] = p1.reshape(s, P)
                h["w"][n:n + s] = w
                h["ls"][n:n + s] = index_pairs[:, 0]
                h["le"][n:n + s] = index_pairs[:, 1]
                n += s
            lens.append(s)
            if progress is not None:
                progress.update(task, advance=1)
            if n >= chunk:
                flush()
        flush()
        errors.check()
        return dense

    def organize_fluxMatrix(self, use_ray=False, progress_bar=None, **args):
        """reference: _fluxmatrix.py:347-415.  Dispatch to the cleaning pass of the clustering method; only the
        stratified one exists on this path (``organize_aggregated`` raises DeprecationWarning in the reference)."""
        if not hasattr(self, "clustering_method") or self.clustering_method is None:
            log.warning("self.clustering_method is not set. This may be a model saved before stratified was implemented, "
                        "or you may not have run cluster_coordinates! Assuming the former and setting to aggregated.")
            self.clustering_method = "aggregated"
        if self.clustering_method == "stratified":
            self.organize_stratified(use_ray, progress_bar)
        elif self.clustering_method == "aggregated":
            raise NotImplementedError("msm_we_b200 implements the stratified path only (organize_aggregated is deprecated "
                                      "in the reference, _fluxmatrix.py:452)")
        else:
            raise Exception(f"Unrecognized clustering_method (Had: {self.clustering_method})")

    def get_iter_fluxMatrix(self, n_iter):
        """reference: _fluxmatrix.py:21-72.  Dense ``(n_clusters+2)^2`` ndarray for one iteration."""
        return _to_host(self._flux_device([n_iter]))

    def get_fluxMatrix(self, n_lag, first_iter=1, last_iter=None, iters_to_use=None, use_ray=False,
                       result_batch_size=5, progress_bar=None):
        """reference: _fluxmatrix.py:166-345."""
        from .. import ops

        self._fluxMatrixParams = [n_lag, first_iter, last_iter, iters_to_use]
        if iters_to_use is not None:
            log.debug("Specific iterations to use were provided for fluxmatrix calculation, using those.")
        else:
            if last_iter is None:
                last_iter = self.maxIter
            iters_to_use = range(first_iter + 1, last_iter)

        self.n_lag = n_lag
        self.errorWeight = 0.0
        self.errorCount = 0
        iters_to_use = list(iters_to_use)
        with ProgressBar(progress_bar) as progress:
            task = progress.add_task(description="Constructing flux matrix", total=len(iters_to_use))
            dense = self._flux_device(iters_to_use, progress, task)
        nI = len(iters_to_use)
        if nI == 0:
            # the reference divides zeros by 0 here (numpy warning, NaN matrix)
            with np.errstate(invalid="ignore", divide="ignore"):
                self.fluxMatrixRaw = dense.cpu().numpy() / nI
            return
        ops.divide_(dense, float(nI))
        self.fluxMatrixRaw = None            # (lets the page-locked buffer of the previous result be reused)
        self.fluxMatrixRaw = _to_host(dense)
