"""GPU-resident state and launch sequences of the hot path.

``DeviceClusters`` is the device snapshot of a ``StratifiedClusters`` object (centres of every WE bin
concatenated, per-bin offsets, ``we_remap``, bin mapper, basis/target bounds).  Its methods enqueue the
K0/K1/K3 kernels on the current CUDA stream and return device tensors; host<->device traffic is the
caller's business (the modelWE mixins stage through pinned buffers, bench.py keeps inputs resident).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from .binning import mapper_kind


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("msm_we_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


class DeviceClusters:
    def __init__(self, bin_mapper, centers_per_bin, we_remap, basis_bounds, target_bounds, pcoord_ndim, device=None):
        """centers_per_bin: list over WE bins of ``[K_b, D]`` float64 arrays or ``None`` (unfitted bin)."""
        self.device = require_cuda(device)
        dev = self.device
        self.nbins = len(centers_per_bin)
        sizes = [0 if c is None else int(len(c)) for c in centers_per_bin]
        fitted = [np.asarray(c, dtype=np.float64) for c in centers_per_bin if c is not None and len(c) > 0]
        if not fitted:
            raise AssertionError("no WE bin has cluster centers")
        self.D = int(fitted[0].shape[1])
        for c in fitted:
            if c.shape[1] != self.D:
                raise ValueError("all cluster models must share one feature dimension")
        self.sizes = np.array(sizes, dtype=np.int64)
        self.total = int(self.sizes.sum())          # T: basis label (stratified_clustering.py:143-150)
        self.max_k = int(self.sizes.max())
        offs = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.bin_offset_host = offs
        self.bin_offset = torch.from_numpy(offs).to(dev)
        self.centers = torch.from_numpy(np.ascontiguousarray(np.concatenate(fitted, axis=0))).to(dev)
        self.csq = ops.centers_sqnorm(self.centers)
        remap = np.array([int(we_remap[b]) for b in range(self.nbins)], dtype=np.int32)
        self.we_remap_host = remap
        self.we_remap = torch.from_numpy(remap).to(dev)
        self.pcoord_ndim = int(pcoord_ndim)
        self.basis = np.ascontiguousarray(basis_bounds, dtype=np.float64).reshape(-1, 2)
        self.target = np.ascontiguousarray(target_bounds, dtype=np.float64).reshape(-1, 2)
        self.bin_mapper = bin_mapper
        kind = mapper_kind(bin_mapper)
        if kind == "rectilinear":
            self.mapper = ops.MapperSpec.rectilinear(bin_mapper.boundaries, dev)
        elif kind == "voronoi":
            self.mapper = ops.MapperSpec.voronoi(bin_mapper.centers, dev)
        else:
            self.mapper = ops.MapperSpec.precomputed(self.nbins)
        if self.mapper.nbins != self.nbins:
            raise ValueError(f"bin mapper has {self.mapper.nbins} bins but {self.nbins} cluster models were given")
        self.errors = ops.DeviceErrors(dev)

    # -- K0 ------------------------------------------------------------------------------------
    def bins_and_flags(self, pcoord_dev, pcoord_host=None):
        """(remapped WE bin int32 [N], flag uint8 [N]) for pcoord [N, P] (device)."""
        bin_out = None
        if self.mapper.kind == _lib.MAPPER_PRECOMPUTED:
            if pcoord_host is None:
                pcoord_host = pcoord_dev.cpu().numpy()
            raw = np.asarray(self.bin_mapper.assign(pcoord_host), dtype=np.int32)
            bin_out = torch.from_numpy(raw).to(self.device)
        return ops.bin_flags(pcoord_dev, self.mapper, self.basis, self.target, we_remap=self.we_remap,
                             errors=self.errors, bin_out=bin_out)

    # -- K0 + K1 -------------------------------------------------------------------------------
    def predict(self, X_dev, pcoord_dev, pcoord_host=None, path=_lib.ASSIGN_AUTO):
        """Labels in the ``StratifiedClusters.predict`` convention (basis -> T, target -> T+1)."""
        if X_dev.shape[1] != self.D:
            raise ValueError(f"coordinates have {X_dev.shape[1]} features, cluster centers have {self.D}")
        bins, flags = self.bins_and_flags(pcoord_dev, pcoord_host)
        labels = ops.assign_stratified(X_dev, bins, flags, self.centers, self.csq, self.bin_offset, self.max_k,
                                       path=path, errors=self.errors)
        return labels, bins, flags

    # -- K3 ------------------------------------------------------------------------------------
    def flux(self, parent_labels, child_labels, flag0, flag1, weights, n_clusters, iter_offsets=None, dense=None):
        return ops.flux_accumulate(parent_labels, child_labels, weights, n_clusters, flag0=flag0, flag1=flag1,
                                   iter_offsets=iter_offsets, dense=dense, errors=self.errors)

    # -- K0 + K1 + K3 in one C call -----------------------------------------------------------
    def hotpath_step(self, X2, pcoord2, weights, n_clusters, iter_offsets=None, dense=None, divisor=0.0,
                     labels_out=None, path=_lib.ASSIGN_AUTO):
        """Stacked batch (parents then children): labels of both halves and, when ``dense`` is given,
        the batch's transitions added into it (then divided by ``divisor`` if it is neither 0 nor 1)."""
        if self.mapper.kind == _lib.MAPPER_PRECOMPUTED:
            raise ValueError("hotpath_step needs a mapper the GPU evaluates (rectilinear / Euclidean Voronoi)")
        n2, D = X2.shape
        n = n2 // 2
        P = pcoord2.shape[1]
        if labels_out is None:
            labels_out = torch.empty(n2, dtype=torch.int64, device=self.device)
        nbytes = _lib.lib.mwe_hotpath_workspace_bytes_ex(n, self.nbins, D, self.max_k, int(path))
        ws = ops.Workspace.get(self.device, nbytes)
        n_iters = 0 if iter_offsets is None else iter_offsets.numel() - 1
        lens_p = self.mapper.lens.ctypes.data if self.mapper.lens is not None else None
        _lib.check(_lib.lib.mwe_hotpath_step_f64(
            X2.data_ptr(), X2.stride(0), D, pcoord2.data_ptr(), P, None if weights is None else weights.data_ptr(), n,
            None if iter_offsets is None else iter_offsets.data_ptr(), n_iters, self.mapper.kind, self.mapper.data.data_ptr(),
            lens_p, self.nbins, self.basis.ctypes.data, self.target.ctypes.data, self.we_remap.data_ptr(),
            self.centers.data_ptr(), self.csq.data_ptr(), self.bin_offset.data_ptr(), self.max_k, int(path),
            int(n_clusters), float(divisor), labels_out.data_ptr(), None if dense is None else dense.data_ptr(),
            ws.data_ptr(), ws.numel(), self.errors.counts.data_ptr(), torch.cuda.current_stream().cuda_stream),
            "mwe_hotpath_step_f64")
        return labels_out

    def check_errors(self):
        self.errors.check()
