"""Size-independent properties at the other BASELINE shapes, on slices that fit a test (the oracle is too slow there):

* config 5 (8,000 segments x 256 features, 100 bins x 100 clusters; 60 iterations = 480,000 frames) and config 3
  (4,000 segments x 3,000 features, 50 bins x 50 clusters; 6 iterations): the tcgen05 path labels exactly like the fp64
  path, labels stay inside their WE bin's cluster range, do not depend on how the batch is cut, and a sample equals the
  oracle's; one Lloyd M step conserves weight and mass (sum of the cluster sums == sum of the rows), re-using the
  bucketing of the previous call changes nothing, and an assignment against the means it just produced can only lower
  the k-means objective;
* config 4 (flux stress: 20,000 clusters x 2 colours, M = 40,004; 2^22 weighted transitions): the sorted-COO output is
  sorted, duplicate-free and in range, its values sum to the total weight (1e-12), unit weights give exact integer
  counts, and the COO outputs of two iteration blocks merge (block order = iteration order) into the one-shot result.
"""
import dataclasses

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _shape(name, iters):
    import torch

    import workloads
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.engine import DeviceClusters

    cfg = dataclasses.replace(workloads.CONFIGS[name], n_iters=iters)
    dev = torch.device("cuda:0")
    means, centers = workloads.make_centers(cfg)
    basis, target = workloads.region_bounds(cfg)
    eng = DeviceClusters(RectilinearBinMapper(workloads.boundaries(cfg)), centers, {b: b for b in range(cfg.n_bins)}, basis,
                         target, 1, device=dev)
    return cfg, eng, workloads.generate_device(cfg, dev, means=means), centers


@pytest.mark.parametrize("name,iters", [("cfg5", 60), ("cfg3", 6)])
def test_assignment_properties_at_the_large_shapes(name, iters):
    import torch

    from msm_we_b200 import _lib

    cfg, eng, data, centers = _shape(name, iters)
    X, pc = data["X"], data["pcoord"]
    tc, bins, flags = eng.predict(X, pc, path=_lib.ASSIGN_TF32X3)
    f64, _, _ = eng.predict(X, pc, path=_lib.ASSIGN_FP64)
    assert torch.equal(tc, f64)
    eng.check_errors()
    # inside the WE bin's cluster range (basis -> T, target -> T + 1)
    offs = eng.bin_offset
    T = int(offs[-1])
    free = flags == 0
    b = bins[free].long()
    assert bool(((tc[free] >= offs[b]) & (tc[free] < offs[b + 1])).all())
    assert bool((tc[(flags & 2) != 0] == T + 1).all()) and bool((tc[((flags & 1) != 0) & ((flags & 2) == 0)] == T).all())
    # partition invariance (a cut that is not a multiple of any tile size)
    cut = X.shape[0] // 3 + 77
    a, _, _ = eng.predict(X[:cut], pc[:cut], path=_lib.ASSIGN_TF32X3)
    c, _, _ = eng.predict(X[cut:], pc[cut:], path=_lib.ASSIGN_TF32X3)
    assert torch.equal(tc, torch.cat([a, c]))
    # sampled oracle comparison
    idx = torch.arange(0, X.shape[0], max(1, X.shape[0] // 300), device=X.device)
    Xs, bs, fs, ls = X[idx].cpu().numpy(), bins[idx].cpu().numpy(), flags[idx].cpu().numpy(), tc[idx].cpu().numpy()
    offs_h = offs.cpu().numpy()
    for i in range(len(ls)):
        if fs[i] == 0:
            lab, amb = O.kmeans_assign_tiebreak(Xs[i:i + 1], centers[bs[i]], return_ambiguous=True)
            assert amb[0] or ls[i] == offs_h[bs[i]] + lab[0]


@pytest.mark.parametrize("name,iters", [("cfg5", 60), ("cfg3", 6)])
def test_lloyd_step_properties_at_the_large_shapes(name, iters):
    import torch

    from msm_we_b200 import _lib, ops

    cfg, eng, data, _ = _shape(name, iters)
    N = data["n"]
    Xc, pc0, w = data["X"][N:], data["pcoord"][:N], data["weights"]
    bins, flags = eng.bins_and_flags(pc0)
    centers = eng.centers.clone()
    sumK, D = centers.shape
    path = _lib.ASSIGN_TF32X3
    ws = ops.assign_workspace(Xc, cfg.n_bins, eng.max_k, path)
    labels = torch.empty(N, dtype=torch.int64, device=Xc.device)
    ops.assign_stratified(Xc, bins, flags, centers, ops.centers_sqnorm(centers), eng.bin_offset, eng.max_k, path=path,
                          label_out=labels, workspace=ws, errors=eng.errors)
    first = labels.clone()
    used = flags == 0

    def objective(lab, c):
        d = Xc[used] - c[lab[used]]
        return float((w[used] * (d * d).sum(dim=1)).sum())

    sum_wx, sum_w = ops.centroid_accumulate(Xc, w, labels, sumK)
    # weight and mass conservation over the frames that take part (basis / target parents carry labels >= sumK)
    tw = float(w[used].sum())
    assert abs(float(sum_w.sum()) - tw) <= 1e-12 * tw
    mass = (Xc[used] * w[used, None]).sum(dim=0)
    assert torch.allclose(sum_wx.sum(dim=0), mass, rtol=1e-9, atol=1e-9 * float(mass.abs().max()))
    before = objective(labels, centers)
    ops.lloyd_finalize(sum_wx, sum_w, centers)
    assert objective(labels, centers) <= before * (1 + 1e-12)             # the mean minimises the weighted objective
    # E step against the new centres, bucketing re-used: identical to a fresh call, and no worse an objective
    ops.assign_stratified(Xc, bins, flags, centers, ops.centers_sqnorm(centers), eng.bin_offset, eng.max_k, path=path,
                          label_out=labels, workspace=ws, reuse_buckets=True, errors=eng.errors)
    fresh = ops.assign_stratified(Xc, bins, flags, centers, ops.centers_sqnorm(centers), eng.bin_offset, eng.max_k, path=path,
                                  errors=eng.errors)
    assert torch.equal(labels, fresh)
    assert objective(labels, centers) <= objective(first, centers) * (1 + 1e-12)
    eng.check_errors()


def test_flux_stress_coo_properties():
    import torch

    import workloads
    from msm_we_b200 import ops

    dev = torch.device("cuda:0")
    d = workloads.generate_cfg4_device(dev, 1 << 22)
    n, N, offs = d["n_clusters"], d["n"], d["iter_offsets"]
    M = 2 * (n + 2)

    def coo(w, lo=0, hi=None):
        hi = N if hi is None else hi
        i0 = int(torch.searchsorted(offs, torch.tensor(lo, device=dev)))
        i1 = int(torch.searchsorted(offs, torch.tensor(hi, device=dev)))
        sub = (offs[i0:i1 + 1] - offs[i0]).contiguous()
        _, (r_, c_, v_, nnz) = ops.flux_accumulate(d["start"][lo:hi].contiguous(), d["end"][lo:hi].contiguous(),
                                                   None if w is None else w[lo:hi].contiguous(), n,
                                                   col0=d["col0"][lo:hi].contiguous(), col1=d["col1"][lo:hi].contiguous(), C=2,
                                                   iter_offsets=sub, want_coo=True, dense=None)
        k = int(nnz.item())
        return r_[:k], c_[:k], v_[:k]

    rows, cols, vals = coo(d["w"])
    r, c, v = rows.cpu().numpy(), cols.cpu().numpy(), vals.cpu().numpy()
    assert r.min() >= 0 and r.max() < M and c.min() >= 0 and c.max() < M
    key = r.astype(np.int64) * M + c
    assert (np.diff(key) > 0).all()                                      # sorted by cell, no duplicates
    total = float(d["w"].sum())
    assert abs(v.sum() - total) <= 1e-12 * total
    # unit weights: exact transition counts
    r1, c1, v1 = coo(None)
    v1 = v1.cpu().numpy()
    assert np.array_equal(v1, np.rint(v1)) and v1.sum() == N
    assert torch.equal(r1, rows) and torch.equal(c1, cols)
    # two iteration blocks: the union of the blocks' cells is the whole's cell set and the values add up per cell
    cut = int(offs[(offs.numel() - 1) // 2])
    ra, ca, va = coo(d["w"], 0, cut)
    rb, cb, vb = coo(d["w"], cut, N)
    ka = (ra * M + ca).cpu().numpy(); kb = (rb * M + cb).cpu().numpy()
    merged = dict(zip(ka.tolist(), va.cpu().numpy().tolist()))
    for k_, x in zip(kb.tolist(), vb.cpu().numpy().tolist()):
        merged[k_] = merged.get(k_, 0.0) + x                              # block order = iteration order
    assert sorted(merged) == key.tolist()
    got = np.array([merged[k_] for k_ in key.tolist()])
    assert np.allclose(got, v, rtol=1e-12, atol=0)
