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
    sumK, D = centers_dev.shape
    sums = torch.empty(sumK * (D + 1), dtype=torch.float64, device=centers_dev.device)   # sum_wx | sum_w: one exchange
    offs_host = bin_offset_dev.cpu().numpy() if relocate_empty else None
    # the points and their WE bins do not change between iterations: K1 buckets them once (private workspace + one
    # label buffer, so nothing else writes to them in between)
    nbins = bin_offset_dev.numel() - 1
    k1_ws = ops.assign_workspace(X_dev, nbins, max_k, path) if n_iter > 1 else None
    labels = torch.empty(X_dev.shape[0], dtype=torch.int64, device=X_dev.device) if n_iter > 0 else None
    for it in range(n_iter):
        ops.assign_stratified(X_dev, bins_dev, flags_dev, centers_dev, ops.centers_sqnorm(centers_dev), bin_offset_dev, max_k,
                              path=path, errors=errors, label_out=labels, workspace=k1_ws, reuse_buckets=it > 0)
        sum_wx, sum_w = ops.centroid_accumulate(X_dev, w_dev, labels, sumK, out=sums)
        if group is not None:
            import torch.distributed as dist

            dist.all_reduce(sums, group=group)
        if relocate_empty:
            _relocate_empty_clusters(X_dev, w_dev, labels, centers_dev, bins_dev, flags_dev, bin_offset_dev, sum_wx, sum_w, group,
                                     offs_host=offs_host)
        ops.lloyd_finalize(sum_wx, sum_w, centers_dev)
    return labels


# Above this many points in the bins that lost several clusters at once, the farthest points are ranked on the device
# (see _relocate_empty_clusters); below it the host runs the very numpy call sklearn uses.
EXACT_ORDER_MAX_POINTS = 1 << 16


def _relocate_empty_clusters(X_dev, w_dev, labels, centers_dev, bins_dev, flags_dev, bin_offset_dev, sum_wx, sum_w, group,
                             offs_host=None):
    """sklearn's ``_relocate_empty_clusters_dense`` per WE-bin model, applied to the (all-reduced) partial sums before
    the mean: a cluster that received no weight takes the point farthest from its own centre, which leaves its old
    cluster.

    Host round trips are what this costs, so there are few: (1) the [sumK] weight sums (every Lloyd iteration; nothing
    else happens when no fitted bin owns an empty cluster); when one does: (2) the indices of the affected bins' points,
    whose distances to their centres are formed on the device (``point_center_dist2``); (3) those distances, from which
    the host picks each bin's farthest points with the numpy call sklearn uses; (4) the winners' old labels.  The rows
    themselves never leave the device: they are gathered, (with a process group: all-gathered as one small tensor, the
    global winners picked from every rank's candidates,) and subtracted / assigned by indexed device updates."""
    import os, time
    dbg = os.environ.get("MWE_RELOC_DEBUG")
    marks = []

    def mark(tag):
        if dbg:
            torch.cuda.synchronize() if X_dev.is_cuda else None
            marks.append((tag, time.perf_counter()))

    mark("start")
    sw = sum_w.cpu().numpy()                                                     # (1)
    if (sw != 0).all():
        return False
    offs = bin_offset_dev.cpu().numpy() if offs_host is None else offs_host
    nbins = len(offs) - 1
    # bins with a model (>= 1 cluster) that received weight and own an empty cluster; the non-empty cluster ranges tile
    # [0, sumK), so one reduceat per quantity covers them (weights are non-negative: the sum is 0 iff every entry is)
    fitted = np.flatnonzero(np.diff(offs) > 0)
    if fitted.size == 0:
        return False
    first = offs[fitted]
    has_empty = np.add.reduceat((sw == 0).astype(np.int64), first) > 0
    has_weight = np.add.reduceat(sw, first) != 0
    affected = fitted[has_empty & has_weight].tolist()
    if not affected:          # (a model without any point is not being fitted at all)
        return False
    world, rank = 1, 0
    if group is not None:
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = X_dev.device
    D = X_dev.shape[1]
    mask = torch.zeros(nbins, dtype=torch.bool, device=dev)
    mask[torch.tensor(affected, device=dev)] = True
    sel = mask[bins_dev]
    if flags_dev is not None:
        sel &= flags_dev == 0
    mark("host-affected")
    idx = torch.nonzero(sel).squeeze(1).to(torch.int32)                          # (2)  ascending point index
    mark("nonzero")
    d2_dev = ops.point_center_dist2(X_dev, idx, labels, centers_dev)
    mark("dist2")
    bins_aff = bins_dev[idx.long()].to(torch.int64)
    # group the listed points by WE bin (stable: ascending point index inside a bin, the row order sklearn sees) and take
    # every bin's largest distance on the device; only bins that lost SEVERAL clusters need their whole distance list
    members, seg_start = ops.group_by_label(bins_aff, nbins)
    _, _, _, vmax = ops.label_stats(d2_dev, members, seg_start, nbins)
    mark("group+stats")
    at_max = torch.nonzero(d2_dev == vmax[bins_aff]).squeeze(1)                  # (3)  ~ one position per affected bin
    packed = torch.stack([at_max.to(torch.float64), bins_aff[at_max].to(torch.float64), d2_dev[at_max]]).cpu().numpy()
    at_max_h, at_max_bin, d2_at_max = packed[0].astype(np.int64), packed[1].astype(np.int64), packed[2]
    seg_h = seg_start.cpu().numpy()
    mark("argmax-d2h")
    # slots: one per empty cluster of an affected bin, in ascending cluster index (= grouped by bin); every rank sees the
    # same sums, hence the same slots.  Vectorised: a Python loop only over the bins that lost SEVERAL clusters.
    sizes = np.diff(offs)
    bin_of_cluster = np.repeat(np.arange(nbins), sizes)
    aff_mask = np.zeros(nbins, dtype=bool)
    aff_mask[affected] = True
    slot_new = np.flatnonzero((sw == 0) & aff_mask[bin_of_cluster])
    slot_bin = bin_of_cluster[slot_new]
    E = len(slot_new)
    n_empty = np.bincount(slot_bin, minlength=nbins)
    n_here = np.diff(seg_h[: nbins + 1])
    local_pos = np.full(E, -1, dtype=np.int64)
    local_d2 = np.full(E, -np.inf)
    # single-empty bins: the bin's farthest point (first position holding the bin maximum)
    ub, first = np.unique(at_max_bin, return_index=True)
    best_pos = np.full(nbins, -1, dtype=np.int64)
    best_d2 = np.full(nbins, -np.inf)
    best_pos[ub], best_d2[ub] = at_max_h[first], d2_at_max[first]
    single = (n_empty[slot_bin] == 1) & (n_here[slot_bin] > 0)
    local_pos[single] = best_pos[slot_bin[single]]
    local_d2[single] = best_d2[slot_bin[single]]
    # bins that lost several clusters: which far point goes to which cluster follows numpy's argpartition order in
    # sklearn, so the same call runs on each such bin's full distance list -- fetched for all of them in one transfer
    multi = np.flatnonzero((n_empty > 1) & (n_here > 0))
    if multi.size:
        starts, ends = seg_h[multi].astype(np.int64), seg_h[multi + 1].astype(np.int64)
        lens = ends - starts
        slot_first = np.searchsorted(slot_bin, multi)                 # first slot of each multi-empty bin
        if int(lens.sum()) <= EXACT_ORDER_MAX_POINTS:
            gather = torch.from_numpy(np.concatenate([np.arange(a, z) for a, z in zip(starts, ends)])).to(dev)
            rows_all = members[gather].long()
            packed = torch.stack([d2_dev[rows_all], rows_all.to(torch.float64)]).cpu().numpy()
            multi_d2, multi_rows = packed[0], packed[1].astype(np.int64)
            cuts = np.concatenate([[0], np.cumsum(lens)])
            for i, b in enumerate(multi):
                dist2, rows = multi_d2[cuts[i]:cuts[i + 1]], multi_rows[cuts[i]:cuts[i + 1]]
                take = min(int(n_empty[b]), len(dist2))
                far = np.argpartition(dist2, -take)[:-take - 1:-1]
                local_pos[slot_first[i]:slot_first[i] + take] = rows[far]
                local_d2[slot_first[i]:slot_first[i] + take] = dist2[far]
        else:
            # large models: shipping whole distance lists to the host would dominate the Lloyd iteration.  Each such
            # bin's n_empty farthest points are picked on the device (largest distance first, equal distances in point
            # order -- the head of a stable descending sort) and only they come back.  Same SET of relocated points as
            # sklearn; they are handed to the bin's empty clusters in descending distance (sklearn: in the order
            # numpy's introselect leaves them), so the relocated centres may sit at permuted cluster indices of that bin.
            take = np.minimum(n_empty[multi], lens)
            kmax = int(take.max())
            if kmax <= ops.SEGMENT_TOPK_MAX:
                pos_d, val_d = ops.segment_topk(d2_dev, members, seg_start, torch.from_numpy(multi.astype(np.int32)).to(dev), kmax)
                packed = torch.stack([pos_d.to(torch.float64), val_d]).cpu().numpy()          # [2, n_multi, kmax]
                for i in range(len(multi)):
                    local_pos[slot_first[i]:slot_first[i] + take[i]] = packed[0, i, :take[i]].astype(np.int64)
                    local_d2[slot_first[i]:slot_first[i] + take[i]] = packed[1, i, :take[i]]
            else:
                # (more than 8 clusters of one bin emptied at once: rank everything with two stable sorts)
                by_d = torch.argsort(d2_dev, descending=True, stable=True)
                by_bin = torch.argsort(bins_aff[by_d], stable=True)
                ranked = by_d[by_bin]                                          # positions grouped by bin, farthest first
                want = np.concatenate([np.arange(a, a + t) for a, t in zip(starts, take)])
                sel_pos = ranked[torch.from_numpy(want).to(dev)]
                packed = torch.stack([sel_pos.to(torch.float64), d2_dev[sel_pos]]).cpu().numpy()
                cuts = np.concatenate([[0], np.cumsum(take)])
                for i in range(len(multi)):
                    local_pos[slot_first[i]:slot_first[i] + take[i]] = packed[0][cuts[i]:cuts[i + 1]].astype(np.int64)
                    local_d2[slot_first[i]:slot_first[i] + take[i]] = packed[1][cuts[i]:cuts[i + 1]]
    mark(f"host-pick({len(affected)} bins, {E} slots, {idx.numel()} pts)")
    pos_t = torch.from_numpy(np.maximum(local_pos, 0)).to(dev)
    pts = idx.long()[pos_t] if idx.numel() else torch.zeros(E, dtype=torch.int64, device=dev)
    cand = torch.empty((E, D + 3), dtype=torch.float64, device=dev)             # row | weight | old label | distance
    cand[:, :D] = X_dev[pts]
    cand[:, D] = 1.0 if w_dev is None else w_dev[pts]
    cand[:, D + 1] = labels[pts].to(torch.float64)
    cand[:, D + 2] = torch.from_numpy(local_d2).to(dev)
    if world > 1:
        import torch.distributed as dist

        parts = [torch.empty((E, D + 3), dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(parts, cand.contiguous(), group=group)
        cand = torch.cat(parts, dim=0)
    meta = cand[:, D:].cpu().numpy()                                             # (4)  [world*E, 3]
    # winners: per bin the n_empty largest distances among all ranks' candidates (rank order breaks ties)
    dmat = meta[:, 2].reshape(world, E)
    win_rank = np.argmax(dmat, axis=0)                                          # single-empty bins: the best rank's candidate
    win_rows = win_rank * E + np.arange(E)
    valid = np.isfinite(dmat[win_rank, np.arange(E)])
    if world > 1:
        for b in np.flatnonzero(n_empty > 1):                                   # several empties: global top-k of the bin
            slots = np.flatnonzero(slot_bin == b)
            rows_b = (np.arange(world)[:, None] * E + slots[None, :]).ravel()
            d_b = meta[rows_b, 2]
            order = np.argsort(-d_b, kind="stable")[: len(slots)]
            win_rows[slots] = rows_b[order]
            valid[slots] = np.isfinite(d_b[order])
    win_rows, win_new = win_rows[valid], slot_new[valid]
    if win_rows.size == 0:
        return False
    wr = torch.from_numpy(win_rows).to(dev)
    new_ids = torch.from_numpy(win_new).to(dev)
    wts = cand[wr, D]
    delta = cand[wr, :D] * wts[:, None]
    old_h = meta[win_rows, 1].astype(np.int64)
    old_ids = torch.from_numpy(old_h).to(dev)
    # one old cluster may lose several points: subtract them in rounds (the r-th loss of every cluster in round r), so
    # each round's indexed update touches distinct rows and the order of the subtractions is fixed
    order = np.argsort(old_h, kind="stable")
    rank = np.empty(len(old_h), dtype=np.int64)
    sorted_old = old_h[order]
    first = np.r_[0, np.flatnonzero(np.diff(sorted_old)) + 1]
    rank[order] = np.arange(len(old_h)) - np.repeat(first, np.diff(np.r_[first, len(old_h)]))
    for r in range(int(rank.max()) + 1):
        sel_r = torch.from_numpy(np.flatnonzero(rank == r)).to(dev)
        sum_wx.index_add_(0, old_ids[sel_r], -delta[sel_r])
        sum_w.index_add_(0, old_ids[sel_r], -wts[sel_r])
    sum_wx[new_ids] = delta
    sum_w[new_ids] = wts
    mark("apply")
    if dbg:
        print("relocate: " + "  ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}ms" for a, b in zip(marks[:-1], marks[1:])), flush=True)
    return True
