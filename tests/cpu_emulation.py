"""TEST INFRASTRUCTURE: numpy / oracle stand-ins for the C-ABI kernels, so the HOST logic of the drop-in layer
(batch planning, staging, iteration bookkeeping, cleaning, block validation ...) can run in the CPU-only container
against the reference-executed fixtures.  Nothing here ships: ``emulate_kernels(monkeypatch)`` patches
``msm_we_b200.ops`` for the duration of one ``-m "not gpu"`` test; the ``-m gpu`` tests run the same checks through
the real library.  Every stand-in follows the contract documented in include/msm_we_b200.h.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import oracle as O


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a, **k):
        pass

    def synchronize(self):
        pass


class _Stream:
    cuda_stream = 0


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def emulate_kernels(monkeypatch):
    import msm_we_b200.engine as engine
    import msm_we_b200.ops as ops
    import msm_we_b200.clustering_ops as cops
    from msm_we_b200 import _lib

    cpu = torch.device("cpu")
    monkeypatch.setattr(engine, "require_cuda", lambda device=None: cpu)
    monkeypatch.setattr(cops, "require_cuda", lambda device=None: cpu)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a, **k: (64 << 30, 64 << 30))
    real_empty = torch.empty

    def empty(*a, pin_memory=False, **k):
        return real_empty(*a, **k)

    monkeypatch.setattr(torch, "empty", empty)

    def centers_sqnorm(centers):
        c = _np(centers)
        return torch.from_numpy(np.einsum("ij,ij->i", c, c))

    def bin_flags(pcoord, mapper, basis_bounds, target_bounds, we_remap=None, errors=None, bin_out=None, bin_count=None):
        pc = _np(pcoord)
        if pc.ndim == 1:
            pc = pc[:, None]
        N = pc.shape[0]
        if mapper.kind == _lib.MAPPER_PRECOMPUTED:
            raw = _np(bin_out).astype(np.int64)
        else:
            data = _np(mapper.data)
            try:
                if mapper.kind == _lib.MAPPER_RECTILINEAR:
                    lens = list(mapper.lens)
                    bs, p = [], 0
                    for ln in lens:
                        bs.append(data[p:p + ln]); p += ln
                    raw = O.RectilinearBinMapperOracle(bs).assign(pc)
                else:
                    raw = O.VoronoiBinMapperOracle(data).assign(pc)
            except ValueError:
                if errors is not None:
                    errors.counts[_lib.ERR_OUT_OF_BINSPACE] += 1
                raw = np.zeros(N, dtype=np.int64)
            bin_out = torch.empty(N, dtype=torch.int32)
        b = raw if we_remap is None else _np(we_remap)[raw]
        bin_out.copy_(torch.from_numpy(np.asarray(b, dtype=np.int32)))
        basis = np.asarray(basis_bounds, dtype=np.float64).reshape(-1, 2)[: pc.shape[1]]
        target = np.asarray(target_bounds, dtype=np.float64).reshape(-1, 2)[: pc.shape[1]]
        flag = (O.is_we_region(pc, basis).astype(np.uint8) | (O.is_we_region(pc, target).astype(np.uint8) << 1))
        return bin_out, torch.from_numpy(flag)

    def assign_workspace(X, nbins, max_k, path=0):
        return torch.empty(16, dtype=torch.uint8)

    def assign_stratified(X, bin, flag, centers, csq, bin_offset, max_k, path=0, want_local=False, errors=None,
                          label_out=None, bin_count=None, workspace=None, reuse_buckets=False):
        assert not reuse_buckets or (workspace is not None and label_out is not None)
        x, b, c, offs = _np(X), _np(bin).astype(np.int64), _np(centers), _np(bin_offset)
        f = np.zeros(len(b), dtype=np.uint8) if flag is None else _np(flag)
        total = int(offs[-1])
        out = np.zeros(len(b), dtype=np.int64)
        local = np.full(len(b), -1, dtype=np.int32)
        out[(f & 2) != 0] = total + 1
        out[((f & 1) != 0) & ((f & 2) == 0)] = total
        free = f == 0
        for wb in np.unique(b[free]):
            sel = np.flatnonzero(free & (b == wb))
            lo, hi = int(offs[wb]), int(offs[wb + 1])
            if hi == lo:
                if errors is not None:
                    errors.counts[_lib.ERR_NO_CENTERS] += len(sel)
                continue
            lab = O.kmeans_assign_tiebreak(x[sel], c[lo:hi])
            out[sel] = lab + lo
            local[sel] = lab
        res = torch.from_numpy(out)
        if label_out is not None:
            label_out.copy_(res)
            res = label_out
        return (res, torch.from_numpy(local)) if want_local else res

    def minibatch_update(X, w, label, centers, counts):
        O.minibatch_update(_np(X), _np(w), centers.numpy(), counts.numpy(), _np(label))

    def centroid_accumulate(X, w, label, sumK, out=None):
        x, lab = _np(X), _np(label)
        ww = np.ones(len(lab)) if w is None else _np(w)
        sums = np.zeros((sumK, x.shape[1]))
        wsum = np.zeros(sumK)
        ok = (lab >= 0) & (lab < sumK)
        for i in np.flatnonzero(ok):                 # sample order, product and sum rounded separately
            wsum[lab[i]] += ww[i]
            sums[lab[i]] = sums[lab[i]] + x[i] * ww[i]
        if out is not None:
            D = x.shape[1]
            out[: sumK * D].copy_(torch.from_numpy(sums).reshape(-1))
            out[sumK * D:].copy_(torch.from_numpy(wsum))
            return out[: sumK * D].view(sumK, D), out[sumK * D:]
        return torch.from_numpy(sums), torch.from_numpy(wsum)

    def lloyd_finalize(sum_wx, sum_w, centers):
        sw = _np(sum_w)
        nz = sw > 0
        c = centers.numpy()
        c[nz] = _np(sum_wx)[nz] * (1.0 / sw[nz])[:, None]

    def flux_accumulate(start, end, w, n_clusters, flag0=None, flag1=None, col0=None, col1=None, C=1, iter_offsets=None,
                        dense=None, want_coo=False, errors=None):
        s, e, ww = _np(start).copy(), _np(end).copy(), _np(w)
        n = int(n_clusters)
        M = C * (n + 2)
        if flag1 is not None:
            f1 = _np(flag1)
            e[(f1 & 2) != 0] = n + 1
        if flag0 is not None:
            s[(_np(flag0) & 1) != 0] = n
        if flag1 is not None:
            e[(f1 & 1) != 0] = n
        bad = (s < 0) | (s >= n + 2) | (e < 0) | (e >= n + 2)
        if bad.any():
            if errors is not None:
                errors.counts[_lib.ERR_LABEL_RANGE] += int(bad.sum())
            s, e, ww = s[~bad], e[~bad], ww[~bad]
        if C == 2:
            s = 2 * s + _np(col0)[~bad]
            e = 2 * e + _np(col1)[~bad]
        if dense is None and not want_coo:
            dense = torch.zeros(M, M, dtype=torch.float64)
        offs = np.array([0, len(ww)]) if iter_offsets is None else _np(iter_offsets)
        if bad.any():
            offs = np.array([0, len(ww)])
        from scipy.sparse import coo_matrix

        total = np.zeros((M, M))
        for a, b in zip(offs[:-1], offs[1:]):
            total = total + np.asarray(coo_matrix((ww[a:b], (s[a:b], e[a:b])), shape=(M, M)).todense())
        if dense is not None:
            d = dense.numpy()
            d += total
        if want_coo:
            r, c = np.nonzero(total)
            nnz = torch.tensor([len(r)], dtype=torch.int64)
            return dense, (torch.from_numpy(r.astype(np.int64)), torch.from_numpy(c.astype(np.int64)),
                           torch.from_numpy(total[r, c]), nnz)
        return dense

    def divide_(buf, divisor):
        buf /= divisor
        return buf

    def group_by_label(labels, n_labels):
        lab = _np(labels)
        key = np.where((lab >= 0) & (lab < n_labels), lab, n_labels)
        order = np.argsort(key, kind="stable").astype(np.int32)
        seg = np.searchsorted(key[order], np.arange(n_labels + 2)).astype(np.int32)
        return torch.from_numpy(order), torch.from_numpy(seg)

    def label_stats(values, members, seg_start, n_labels):
        v, m, s = _np(values), _np(members), _np(seg_start)
        cnt = np.zeros(n_labels, dtype=np.int64)
        out = np.zeros((3, n_labels))
        out[1], out[2] = np.inf, -np.inf
        for k in range(n_labels):
            x = v[m[s[k]:s[k + 1]]]
            x = x[~np.isnan(x)]
            cnt[k] = len(x)
            if len(x):
                out[0, k], out[1, k], out[2, k] = x.sum(), x.min(), x.max()
        return torch.from_numpy(cnt), torch.from_numpy(out[0]), torch.from_numpy(out[1]), torch.from_numpy(out[2])

    def segment_topk(values, members, seg_start, seg_ids, k):
        v, m, s, ids = _np(values), _np(members).astype(np.int64), _np(seg_start), _np(seg_ids)
        pos = np.full((len(ids), k), -1, dtype=np.int32)
        val = np.full((len(ids), k), -np.inf)
        for i, g in enumerate(ids):
            mem = m[s[g]:s[g + 1]]
            mem = mem[~np.isnan(v[mem])]
            order = mem[np.lexsort((mem, -v[mem]))][:k]          # value descending, equal values in member order
            pos[i, :len(order)] = order
            val[i, :len(order)] = v[order]
        return torch.from_numpy(pos), torch.from_numpy(val)

    def rows_with_nan(X):
        return torch.from_numpy(np.isnan(_np(X)).any(axis=1).astype(np.uint8))

    def point_center_dist2(X, index_list, labels, centers):
        x, idx = _np(X), _np(index_list).astype(np.int64)
        c = _np(centers)[_np(labels)[idx]]
        return torch.from_numpy(((x[idx] - c) ** 2).sum(axis=1))

    def project(X, components, mean=None, out=None):
        r = O.linear_transform(_np(X), _np(components), None if mean is None else _np(mean))
        return torch.from_numpy(np.ascontiguousarray(r))

    for name, fn in dict(centers_sqnorm=centers_sqnorm, bin_flags=bin_flags, assign_stratified=assign_stratified, assign_workspace=assign_workspace,
                         minibatch_update=minibatch_update, centroid_accumulate=centroid_accumulate,
                         lloyd_finalize=lloyd_finalize, flux_accumulate=flux_accumulate, divide_=divide_,
                         group_by_label=group_by_label, label_stats=label_stats, segment_topk=segment_topk, rows_with_nan=rows_with_nan, point_center_dist2=point_center_dist2,
                         project=project).items():
        monkeypatch.setattr(ops, name, fn)
