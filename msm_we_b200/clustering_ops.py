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


def lloyd_fit(X_dev, w_dev, bins_dev, centers_dev, bin_offset_dev, max_k, n_iter, group=None, flags_dev=None,
              path=None, errors=None, relocate_empty=True):
    """``n_iter`` full Lloyd iterations of every WE bin's model on device-resident data (BASELINE config 5): per
    iteration ONE K1 launch sequence labels the points of all bins, ONE K2 launch sequence forms every cluster's
    ``sum w x`` / ``sum w`` in sample order, and the mean replaces the centre (clusters without members keep theirs
    unless relocated).  ``centers_dev`` is updated in place; returns the labels of the last E step.

    reference arithmetic: ``KMeans.fit`` -> ``lloyd_iter_chunked_dense(update_centers=True)``
    (sklearn/cluster/_k_means_lloyd.pyx:23-165; reached from msm_we/_hamsm/_clustering.py:289,491), including its
    empty-cluster relocation (``_relocate_empty_clusters_dense``, sklearn/cluster/_k_means_common.pyx): a cluster that
    received no weight takes the point farthest from its own centre, which is removed from its old cluster's sum.

    With a process group every rank holds its own points (iteration-range shard); the partial sums are all-reduced --
    the only exchange step of the clustering path -- and every rank applies the identical finalize."""
    from . import _lib

    if path is None:
        path = _lib.ASSIGN_AUTO
    labels = None
    sumK = centers_dev.shape[0]
    for _ in range(n_iter):
        labels = ops.assign_stratified(X_dev, bins_dev, flags_dev, centers_dev, ops.centers_sqnorm(centers_dev),
                                       bin_offset_dev, max_k, path=path, errors=errors)
        sum_wx, sum_w = ops.centroid_accumulate(X_dev, w_dev, labels, sumK)
        if group is not None:
            import torch.distributed as dist

            dist.all_reduce(sum_wx, group=group)
            dist.all_reduce(sum_w, group=group)
        if relocate_empty:
            _relocate_empty_clusters(X_dev, w_dev, labels, centers_dev, bins_dev, flags_dev, bin_offset_dev, sum_wx, sum_w, group)
        ops.lloyd_finalize(sum_wx, sum_w, centers_dev)
    return labels


def _relocate_empty_clusters(X_dev, w_dev, labels, centers_dev, bins_dev, flags_dev, bin_offset_dev, sum_wx, sum_w, group):
    """sklearn's ``_relocate_empty_clusters_dense`` per WE-bin model, applied to the (all-reduced) partial sums before
    the mean: a cluster that received no weight takes the point farthest from its own centre, which leaves its old
    cluster.  The decision needs one [sumK] read per Lloyd iteration; when a bin does own an empty cluster, the distances
    of THAT bin's points to their centres are formed on the device (``point_center_dist2``), only distances + indices
    come to the host, where the farthest points are picked with the numpy call sklearn uses, and only those few rows
    are fetched."""
    sw = sum_w.cpu().numpy()
    if (sw != 0).all():
        return False
    offs = bin_offset_dev.cpu().numpy()
    nbins = len(offs) - 1
    affected = []
    for b in range(nbins):
        lo, hi = int(offs[b]), int(offs[b + 1])
        if hi > lo and (sw[lo:hi] == 0).any() and sw[lo:hi].sum() != 0:   # a model without any point is not being fitted
            affected.append(b)
    world = 1
    if group is not None:
        import torch.distributed as dist

        world = dist.get_world_size(group)
    if not affected:
        return False
    dev = X_dev.device
    mask = torch.zeros(nbins, dtype=torch.bool, device=dev)
    mask[torch.tensor(affected, device=dev)] = True
    sel = mask[bins_dev.long()]
    if flags_dev is not None:
        sel &= flags_dev == 0
    idx = torch.nonzero(sel).squeeze(1).to(torch.int32)
    d2 = ops.point_center_dist2(X_dev, idx, labels, centers_dev).cpu().numpy()
    idx_h = idx.cpu().numpy()
    bin_h = bins_dev[idx.long()].cpu().numpy()
    changed = False
    for b in affected:
        lo, hi = int(offs[b]), int(offs[b + 1])
        empty = lo + np.flatnonzero(sw[lo:hi] == 0)
        rows = np.flatnonzero(bin_h == b)                 # ascending point index = the row order sklearn sees
        dist2 = d2[rows]
        n_empty = int(empty.size)
        take = min(n_empty, len(dist2))
        far = np.argpartition(dist2, -take)[:-take - 1:-1] if take else np.zeros(0, dtype=np.int64)
        pts = torch.from_numpy(idx_h[rows[far]].astype(np.int64)).to(dev)
        xs = X_dev[pts].cpu().numpy() if take else np.zeros((0, X_dev.shape[1]))
        ws = np.ones(take) if w_dev is None else w_dev[pts].cpu().numpy()
        ls = labels[pts].cpu().numpy() if take else np.zeros(0, dtype=np.int64)
        cand = [(float(dist2[f]), xs[i], float(ws[i]), int(ls[i])) for i, f in enumerate(far)]
        if world > 1:
            import torch.distributed as dist

            gathered = [None] * world
            dist.all_gather_object(gathered, cand, group=group)
            cand = sorted((c for part in gathered for c in part), key=lambda c: -c[0])[:n_empty]
        for new_id, (_, x, wt, old_id) in zip(empty, cand):
            delta = torch.from_numpy(x * wt).to(dev)
            sum_wx[old_id] -= delta
            sum_wx[new_id] = delta
            sum_w[new_id] = wt
            sum_w[old_id] -= wt
            changed = True
    return changed
