"""Thin torch-tensor wrappers over the C ABI (include/msm_we_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every computation is one of
the hand-written sm_100a kernels behind ``libmsm_we_b200.so``.  Nothing in this module computes on
the CPU and nothing falls back to torch ops.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import lib, check


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise TypeError(f"{name}: expected a contiguous tensor")


class Workspace:
    """Grow-only scratch buffer, one per device."""

    _cache = {}

    @classmethod
    def get(cls, device, nbytes: int) -> torch.Tensor:
        key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = None
            cls._cache.pop(key, None)
            buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=f"cuda:{key}")
            cls._cache[key] = buf
        return buf

    @classmethod
    def clear(cls):
        cls._cache.clear()


class DeviceErrors:
    """The device-side error counters of the ABI, turned back into the reference's exceptions."""

    def __init__(self, device):
        self.counts = torch.zeros(_lib.ERR_SLOTS, dtype=torch.int32, device=device)

    def reset(self):
        self.counts.zero_()

    def check(self):
        c = self.counts.cpu().numpy()
        self.counts.zero_()
        if c[_lib.ERR_OUT_OF_BINSPACE]:
            # westpa's rectilinear_assign raises ValueError for a coordinate outside the bin space
            raise ValueError(f"coordinate outside of bin space ({int(c[_lib.ERR_OUT_OF_BINSPACE])} points)")
        if c[_lib.ERR_NO_CENTERS]:
            # msm_we/stratified_clustering.py:187-189
            raise AssertionError(
                f"Not initialized: {int(c[_lib.ERR_NO_CENTERS])} segments fall in a WE bin without cluster centers")
        if c[_lib.ERR_LABEL_RANGE]:
            # scipy.sparse.coo_matrix raises ValueError for an index outside the shape (_fluxmatrix.py:147-158)
            raise ValueError(f"row/column index exceeds matrix dimensions ({int(c[_lib.ERR_LABEL_RANGE])} transitions)")
        if c[_lib.ERR_INTERNAL]:
            raise RuntimeError("internal device error")


class MapperSpec:
    """Device-resident description of a WE bin mapper for K0."""

    def __init__(self, kind, nbins, data=None, lens=None):
        self.kind = kind
        self.nbins = int(nbins)
        self.data = data                      # float32 CUDA tensor or None
        self.lens = None if lens is None else np.ascontiguousarray(lens, dtype=np.int32)

    @classmethod
    def rectilinear(cls, boundaries, device):
        bs = [np.asarray(b, dtype=np.float32) for b in boundaries]
        nbins = int(np.prod([len(b) - 1 for b in bs]))
        data = torch.from_numpy(np.concatenate(bs)).to(device)
        return cls(_lib.MAPPER_RECTILINEAR, nbins, data, [len(b) for b in bs])

    @classmethod
    def voronoi(cls, centers, device):
        c = np.ascontiguousarray(centers, dtype=np.float32)
        if c.ndim == 1:
            c = c[:, None]
        return cls(_lib.MAPPER_VORONOI, c.shape[0], torch.from_numpy(c).to(device))

    @classmethod
    def precomputed(cls, nbins):
        return cls(_lib.MAPPER_PRECOMPUTED, nbins)


def bin_flags(pcoord, mapper: MapperSpec, basis_bounds, target_bounds, we_remap=None, errors: DeviceErrors = None,
              bin_out=None, bin_count=None):
    """K0.  pcoord [N,P] f64 CUDA -> (bin int32 [N], flag uint8 [N])."""
    _req(pcoord, torch.float64, "pcoord")
    if pcoord.dim() == 1:
        pcoord = pcoord[:, None]
    N, P = pcoord.shape
    dev = pcoord.device
    if errors is None:
        errors = DeviceErrors(dev)
    if mapper.kind == _lib.MAPPER_PRECOMPUTED:
        if bin_out is None:
            raise ValueError("precomputed mapper needs bin_out holding the raw bins")
        _req(bin_out, torch.int32, "bin_out")
    elif bin_out is None:
        bin_out = torch.empty(N, dtype=torch.int32, device=dev)
    flag = torch.empty(N, dtype=torch.uint8, device=dev)
    basis = np.ascontiguousarray(basis_bounds, dtype=np.float64).reshape(-1, 2)
    target = np.ascontiguousarray(target_bounds, dtype=np.float64).reshape(-1, 2)
    if basis.shape[0] < P or target.shape[0] < P:
        raise ValueError("basis/target bounds need one (lo, hi) row per pcoord dimension")
    _req(we_remap, torch.int32, "we_remap")
    _req(bin_count, torch.int32, "bin_count")
    lens_p = mapper.lens.ctypes.data if mapper.lens is not None else None
    check(lib.mwe_bin_flags_f64(_ptr(pcoord), N, P, mapper.kind, _ptr(mapper.data), lens_p, mapper.nbins,
                                basis.ctypes.data, target.ctypes.data, _ptr(we_remap), _ptr(bin_out), _ptr(flag),
                                _ptr(bin_count), _ptr(errors.counts), _stream()), "mwe_bin_flags_f64")
    return bin_out, flag


def project(X, components, mean=None, out=None):
    """``(X - mean) @ components.T`` on the device (the reference's ``coordinates.transform`` of a fitted PCA)."""
    if not X.is_cuda or X.dtype != torch.float64:
        raise TypeError("X: expected a float64 CUDA tensor (there is no CPU path)")
    components = components.contiguous()
    _req(components, torch.float64, "components")
    if X.dim() != 2 or components.dim() != 2 or X.shape[1] != components.shape[1] or (X.shape[0] > 1 and X.stride(1) != 1):
        raise ValueError("project: X must be [N, D_in] with unit column stride, components [d_out, D_in]")
    if mean is not None:
        _req(mean, torch.float64, "mean")
        mean = mean.contiguous()
        if mean.numel() != X.shape[1]:
            raise ValueError("project: mean must have D_in entries")
    N, d_out = X.shape[0], components.shape[0]
    if out is None:
        out = torch.empty((N, d_out), dtype=torch.float64, device=X.device)
    check(lib.mwe_project_f64(_ptr(X), N, X.shape[1], X.stride(0), _ptr(components), None if mean is None else _ptr(mean),
                              d_out, _ptr(out), out.stride(0), _stream()), "mwe_project_f64")
    return out


def centers_sqnorm(centers):
    _req(centers, torch.float64, "centers")
    sumK, D = centers.shape
    out = torch.empty(sumK, dtype=torch.float64, device=centers.device)
    check(lib.mwe_centers_sqnorm_f64(_ptr(centers), sumK, D, _ptr(out), _stream()), "mwe_centers_sqnorm_f64")
    return out


def assign_workspace(X, nbins, max_k, path=_lib.ASSIGN_FP64):
    """A private K1 workspace (uint8 tensor) for callers that label the same points repeatedly (``reuse_buckets``)."""
    N, D = X.shape
    return torch.empty(lib.mwe_assign_workspace_bytes_ex(N, int(nbins), D, int(max_k), int(path)), dtype=torch.uint8, device=X.device)


def assign_stratified(X, bin, flag, centers, csq, bin_offset, max_k, path=_lib.ASSIGN_FP64, want_local=False,
                      errors: DeviceErrors = None, label_out=None, bin_count=None, workspace=None, reuse_buckets=False):
    """K1.  X [N,D] f64 (row stride may exceed D) -> labels int64 [N] (and per-bin local argmin).

    ``reuse_buckets``: ``workspace`` (from ``assign_workspace``) and ``label_out`` still hold what the previous call with
    the same X rows, ``bin``, ``flag`` and ``bin_offset`` left there, so the points are not bucketed by WE bin again --
    Lloyd iterations only change the centres."""
    if not X.is_cuda or X.dtype != torch.float64 or X.dim() != 2 or X.stride(1) != 1:
        raise TypeError("X: expected a CUDA float64 [N, D] tensor with unit column stride")
    _req(bin, torch.int32, "bin"); _req(flag, torch.uint8, "flag")
    _req(centers, torch.float64, "centers"); _req(csq, torch.float64, "csq"); _req(bin_offset, torch.int64, "bin_offset")
    N, D = X.shape
    ldx = X.stride(0) if N > 1 else D
    dev = X.device
    nbins = bin_offset.numel() - 1
    if errors is None:
        errors = DeviceErrors(dev)
    if label_out is None:
        label_out = torch.empty(N, dtype=torch.int64, device=dev)
    local = torch.empty(N, dtype=torch.int32, device=dev) if want_local else None
    nbytes = lib.mwe_assign_workspace_bytes_ex(N, nbins, D, int(max_k), int(path))
    if reuse_buckets and (workspace is None or want_local):
        raise ValueError("reuse_buckets needs the private workspace (and label_out) of the previous call")
    if workspace is not None:
        _req(workspace, torch.uint8, "workspace")
        if workspace.numel() < nbytes:
            raise ValueError("workspace too small")
    ws = workspace if workspace is not None else Workspace.get(dev, nbytes)
    check(lib.mwe_assign_stratified_f64(_ptr(X), N, D, ldx, _ptr(bin), _ptr(flag), _ptr(centers), _ptr(csq),
                                        _ptr(bin_offset), nbins, int(max_k),
                                        int(path) | (_lib.ASSIGN_REUSE_BUCKETS if reuse_buckets else 0), _ptr(bin_count), _ptr(label_out),
                                        _ptr(local),
                                        _ptr(ws), ws.numel(), _ptr(errors.counts), _stream()),
          "mwe_assign_stratified_f64")
    return (label_out, local) if want_local else label_out


def _centroid_common(fn, name, X, w, label, sumK, a, b):
    if not X.is_cuda or X.dtype != torch.float64 or X.dim() != 2 or X.stride(1) != 1:
        raise TypeError("X: expected a CUDA float64 [N, D] tensor with unit column stride")
    _req(w, torch.float64, "w"); _req(label, torch.int64, "label")
    N, D = X.shape
    ldx = X.stride(0) if N > 1 else D
    nbytes = lib.mwe_centroid_workspace_bytes(N, sumK)
    ws = Workspace.get(X.device, nbytes)
    check(fn(_ptr(X), N, D, ldx, _ptr(w), _ptr(label), sumK, _ptr(a), _ptr(b), _ptr(ws), ws.numel(), _stream()), name)


def centroid_accumulate(X, w, label, sumK, out=None):
    """K2 partial sums: (sum_wx [sumK,D], sum_w [sumK]).  ``out``: optional flat float64 buffer of sumK*(D+1) elements that
    receives both (sum_wx first), so a multi-GPU caller exchanges them with ONE all-reduce."""
    D = X.shape[1]
    if out is None:
        out = torch.empty(sumK * (D + 1), dtype=torch.float64, device=X.device)
    _req(out, torch.float64, "out")
    if out.numel() != sumK * (D + 1):
        raise ValueError("out must hold sumK * (D + 1) float64 values")
    sum_wx = out[: sumK * D].view(sumK, D)
    sum_w = out[sumK * D:]
    _centroid_common(lib.mwe_centroid_accumulate_f64, "mwe_centroid_accumulate_f64", X, w, label, sumK, sum_wx, sum_w)
    return sum_wx, sum_w


def minibatch_update(X, w, label, centers, counts):
    """K2 fused running-mean update of centers [sumK,D] / counts [sumK], in place."""
    _req(centers, torch.float64, "centers"); _req(counts, torch.float64, "counts")
    _centroid_common(lib.mwe_minibatch_update_f64, "mwe_minibatch_update_f64", X, w, label, centers.shape[0], centers, counts)


def lloyd_finalize(sum_wx, sum_w, centers):
    _req(sum_wx, torch.float64, "sum_wx"); _req(sum_w, torch.float64, "sum_w"); _req(centers, torch.float64, "centers")
    sumK, D = centers.shape
    check(lib.mwe_lloyd_finalize_f64(_ptr(sum_wx), _ptr(sum_w), sumK, D, _ptr(centers), _stream()), "mwe_lloyd_finalize_f64")


def minibatch_finalize(sum_wx, sum_w, centers, counts):
    _req(sum_wx, torch.float64, "sum_wx"); _req(sum_w, torch.float64, "sum_w")
    _req(centers, torch.float64, "centers"); _req(counts, torch.float64, "counts")
    sumK, D = centers.shape
    check(lib.mwe_minibatch_finalize_f64(_ptr(sum_wx), _ptr(sum_w), sumK, D, _ptr(centers), _ptr(counts), _stream()),
          "mwe_minibatch_finalize_f64")


def flux_accumulate(start, end, w, n_clusters, flag0=None, flag1=None, col0=None, col1=None, C=1, iter_offsets=None,
                    dense=None, want_coo=False, errors: DeviceErrors = None):
    """K3.  Adds the transitions into ``dense`` [CM,CM] (created zeroed if None and not want_coo) and/or
    returns sorted COO triples."""
    _req(start, torch.int64, "start"); _req(end, torch.int64, "end"); _req(w, torch.float64, "w")
    for nm, t in (("flag0", flag0), ("flag1", flag1), ("col0", col0), ("col1", col1)):
        _req(t, torch.uint8, nm)
    _req(iter_offsets, torch.int64, "iter_offsets")
    N = start.numel()
    dev = start.device
    CM = C * (n_clusters + 2)
    if errors is None:
        errors = DeviceErrors(dev)
    if dense is None and not want_coo:
        dense = torch.zeros(CM, CM, dtype=torch.float64, device=dev)
    if dense is not None:
        _req(dense, torch.float64, "dense")
        if dense.shape != (CM, CM):
            raise ValueError(f"dense must be [{CM},{CM}]")
    coo_r = coo_c = coo_v = nnz = None
    if want_coo:
        coo_r = torch.empty(max(N, 1), dtype=torch.int64, device=dev)
        coo_c = torch.empty(max(N, 1), dtype=torch.int64, device=dev)
        coo_v = torch.empty(max(N, 1), dtype=torch.float64, device=dev)
        nnz = torch.zeros(1, dtype=torch.int64, device=dev)
    n_iters = 0 if iter_offsets is None else iter_offsets.numel() - 1
    nbytes = lib.mwe_flux_workspace_bytes(N)
    ws = Workspace.get(dev, nbytes)
    check(lib.mwe_flux_accumulate_f64(_ptr(start), _ptr(end), _ptr(flag0), _ptr(flag1), _ptr(col0), _ptr(col1), _ptr(w),
                                      N, int(n_clusters), int(C), _ptr(iter_offsets), n_iters, _ptr(dense), _ptr(coo_r),
                                      _ptr(coo_c), _ptr(coo_v), _ptr(nnz), _ptr(ws), ws.numel(), _ptr(errors.counts),
                                      _stream()), "mwe_flux_accumulate_f64")
    if want_coo:
        return dense, (coo_r, coo_c, coo_v, nnz)
    return dense


def divide_(buf, divisor: float):
    _req(buf, torch.float64, "buf")
    check(lib.mwe_divide_f64(_ptr(buf), buf.numel(), float(divisor), _stream()), "mwe_divide_f64")
    return buf


def sort_pairs_(keys_i64, vals_i32, key_bits: int):
    """In-place stable radix sort (keys reinterpreted as u64, values as u32). Exported for tests."""
    _req(keys_i64, torch.int64, "keys"); _req(vals_i32, torch.int32, "vals")
    N = keys_i64.numel()
    nbytes = lib.mwe_sort_workspace_bytes(N)
    ws = Workspace.get(keys_i64.device, nbytes)
    check(lib.mwe_sort_pairs_u64_u32(_ptr(keys_i64), _ptr(vals_i32), N, int(key_bits), _ptr(ws), ws.numel(), _stream()),
          "mwe_sort_pairs_u64_u32")


def group_by_label(labels, n_labels):
    """Stable grouping of ``labels`` [N] int64: ``(members int32 [N], seg_start int32 [n_labels + 2])`` -- the indices
    of every label's members are ``members[seg_start[l]:seg_start[l + 1]]``, in input order; indices of labels outside
    ``[0, n_labels)`` sit in ``members[seg_start[n_labels]:]``."""
    _req(labels, torch.int64, "labels")
    N = labels.numel()
    dev = labels.device
    members = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    seg_start = torch.empty(int(n_labels) + 2, dtype=torch.int32, device=dev)
    nbytes = lib.mwe_centroid_workspace_bytes(N, int(n_labels))
    ws = Workspace.get(dev, nbytes)
    check(lib.mwe_group_by_label(_ptr(labels), N, int(n_labels), _ptr(members), _ptr(seg_start), _ptr(ws), ws.numel(),
                                 _stream()), "mwe_group_by_label")
    return members[:N], seg_start


def label_stats(values, members, seg_start, n_labels):
    """NaN-skipping per-label ``(count int64, sum, min, max)`` of ``values`` ([N] or one column of an [N, P] tensor)."""
    if not values.is_cuda or values.dtype != torch.float64 or values.dim() != 1:
        raise TypeError("values: expected a 1-D CUDA float64 tensor (a column view is fine)")
    _req(members, torch.int32, "members"); _req(seg_start, torch.int32, "seg_start")
    dev = values.device
    n = int(n_labels)
    count = torch.empty(n, dtype=torch.int64, device=dev)
    out = torch.empty((3, n), dtype=torch.float64, device=dev)
    if values.numel() == 0:                      # no members anywhere: the kernel's answer for an empty label
        count.zero_(); out[0].zero_(); out[1].fill_(float("inf")); out[2].fill_(float("-inf"))
        return count, out[0], out[1], out[2]
    ldv = values.stride(0) if values.numel() > 1 else 1
    check(lib.mwe_label_stats_f64(_ptr(values), ldv, _ptr(members), _ptr(seg_start), n, _ptr(count), _ptr(out[0]),
                                  _ptr(out[1]), _ptr(out[2]), _stream()), "mwe_label_stats_f64")
    return count, out[0], out[1], out[2]


SEGMENT_TOPK_MAX = 8


def segment_topk(values, members, seg_start, seg_ids, k):
    """For each listed group (``seg_ids`` int32, indices into ``seg_start``) the ``k`` (<= 8) largest ``values[member]``:
    ``(pos int32 [n_sel, k], val float64 [n_sel, k])``, largest first, equal values in member order; -1 / -inf padding
    when a group has fewer members."""
    if not values.is_cuda or values.dtype != torch.float64 or values.dim() != 1 or (values.numel() > 1 and values.stride(0) != 1):
        raise TypeError("values: expected a contiguous 1-D CUDA float64 tensor")
    _req(members, torch.int32, "members"); _req(seg_start, torch.int32, "seg_start"); _req(seg_ids, torch.int32, "seg_ids")
    n_sel, k = seg_ids.numel(), int(k)
    pos = torch.full((n_sel, k), -1, dtype=torch.int32, device=values.device)
    val = torch.full((n_sel, k), float("-inf"), dtype=torch.float64, device=values.device)
    if n_sel and values.numel():
        check(lib.mwe_segment_topk_f64(_ptr(values), _ptr(members), _ptr(seg_start), _ptr(seg_ids), n_sel, k, _ptr(pos), _ptr(val),
                                       _stream()), "mwe_segment_topk_f64")
    return pos, val


def rows_with_nan(X):
    """uint8 [N]: 1 where row of X [N, D] float64 holds a NaN."""
    if not X.is_cuda or X.dtype != torch.float64 or X.dim() != 2 or X.stride(1) != 1:
        raise TypeError("X: expected a CUDA float64 [N, D] tensor with unit column stride")
    N, D = X.shape
    out = torch.empty(N, dtype=torch.uint8, device=X.device)
    check(lib.mwe_rows_with_nan_f64(_ptr(X), N, D, X.stride(0) if N > 1 else D, _ptr(out), _stream()), "mwe_rows_with_nan_f64")
    return out


def point_center_dist2(X, index_list, labels, centers):
    """fp64 ``||x_i - centers[labels_i]||^2`` for the points in ``index_list`` (int32 CUDA tensor)."""
    if not X.is_cuda or X.dtype != torch.float64 or X.dim() != 2 or X.stride(1) != 1:
        raise TypeError("X: expected a CUDA float64 [N, D] tensor with unit column stride")
    _req(index_list, torch.int32, "index_list"); _req(labels, torch.int64, "labels"); _req(centers, torch.float64, "centers")
    n = index_list.numel()
    out = torch.empty(n, dtype=torch.float64, device=X.device)
    N, D = X.shape
    check(lib.mwe_point_center_dist2_f64(_ptr(X), X.stride(0) if N > 1 else D, D, _ptr(index_list), n, _ptr(labels),
                                         _ptr(centers), _ptr(out), _stream()), "mwe_point_center_dist2_f64")
    return out
