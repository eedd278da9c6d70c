"""Batched k-means steps over many WE bins at once (the stratified structure is the batch dimension).

reference: the per-bin loop of do_stratified_clustering (msm_we/_hamsm/_clustering.py:890-916) calls
``MiniBatchKMeans.partial_fit`` once per WE bin; sklearn's ``_mini_batch_step``
(sklearn/cluster/_kmeans.py:1566-1684) is labels -> running-mean update -> random reassignment.
Here every bin of the batch goes through ONE K1 launch (labels) and ONE K2 launch (update); the
host keeps only the per-model bookkeeping and the RNG-driven decisions.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .engine import require_cuda


def predict_single(centers, X):
    """``MiniBatchKMeans.predict`` for one model: nearest centre, lowest index on ties (K1)."""
    dev = require_cuda()
    X = np.ascontiguousarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("Expected 2D array")
    c = torch.from_numpy(np.ascontiguousarray(centers, dtype=np.float64)).to(dev)
    if X.shape[1] != c.shape[1]:
        raise ValueError(f"X has {X.shape[1]} features, but the model is expecting {c.shape[1]} features as input.")
    n = X.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int32)
    bins = torch.zeros(n, dtype=torch.int32, device=dev)
    offs = torch.tensor([0, c.shape[0]], dtype=torch.int64, device=dev)
    labels = ops.assign_stratified(torch.from_numpy(X).to(dev), bins, None, c, ops.centers_sqnorm(c), offs, c.shape[0])
    return labels.cpu().numpy().astype(np.int32)


def partial_fit_models(batch, device=None):
    """``batch``: list of ``(model, X_b, sample_weight_or_None)``, one entry per WE bin, processed in
    list order for everything that consumes random numbers."""
    dev = require_cuda(device)
    if not batch:
        return
    prepared = []
    for model, X, w in batch:
        Xc, wc, reassign = model._prepare(X, w)
        prepared.append((model, Xc, wc, reassign))
    D = prepared[0][1].shape[1]
    sizes = [m.cluster_centers_.shape[0] for m, _, _, _ in prepared]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    rows = [x.shape[0] for _, x, _, _ in prepared]
    X_all = torch.from_numpy(np.concatenate([x for _, x, _, _ in prepared], axis=0)).to(dev)
    w_all = torch.from_numpy(np.concatenate([w for _, _, w, _ in prepared])).to(dev)
    bins = torch.from_numpy(np.repeat(np.arange(len(prepared), dtype=np.int32), rows)).to(dev)
    centers = torch.from_numpy(np.concatenate([m.cluster_centers_ for m, _, _, _ in prepared], axis=0)).to(dev)
    counts = torch.from_numpy(np.concatenate([m._counts for m, _, _, _ in prepared])).to(dev)
    if centers.shape[1] != D:
        raise ValueError("feature dimension mismatch between batch and cluster centers")
    errors = ops.DeviceErrors(dev)
    labels = ops.assign_stratified(X_all, bins, None, centers, ops.centers_sqnorm(centers), torch.from_numpy(offs).to(dev),
                                   int(max(sizes)), errors=errors)
    ops.minibatch_update(X_all, w_all, labels, centers, counts)
    centers_h = centers.cpu().numpy()
    counts_h = counts.cpu().numpy()
    errors.check()
    for i, (model, Xc, _, reassign) in enumerate(prepared):
        model.cluster_centers_ = np.ascontiguousarray(centers_h[offs[i]:offs[i + 1]])
        model._counts = np.ascontiguousarray(counts_h[offs[i]:offs[i + 1]])
        model._finish(Xc, reassign)


def lloyd_fit(X_dev, w_dev, bins_dev, centers_dev, bin_offset_dev, max_k, n_iter, group=None):
    """``n_iter`` full Lloyd iterations of every bin's model on device-resident data (BASELINE cfg 5;
    reference arithmetic: sklearn/cluster/_k_means_lloyd.pyx:23-165).  With a process group the partial
    sums are all-reduced, which is the only exchange step of the clustering path.  Returns labels of
    the last E step."""
    labels = None
    sumK = centers_dev.shape[0]
    for _ in range(n_iter):
        labels = ops.assign_stratified(X_dev, bins_dev, None, centers_dev, ops.centers_sqnorm(centers_dev), bin_offset_dev,
                                       max_k)
        sum_wx, sum_w = ops.centroid_accumulate(X_dev, w_dev, labels, sumK)
        if group is not None:
            import torch.distributed as dist

            dist.all_reduce(sum_wx, group=group)
            dist.all_reduce(sum_w, group=group)
        ops.lloyd_finalize(sum_wx, sum_w, centers_dev)
    return labels
