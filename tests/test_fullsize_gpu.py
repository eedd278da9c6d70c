"""Size-independent properties at BASELINE's full cfg2 size (200 WE iterations x 1000 segments x 64 features,
30 bins x 20 clusters): the oracle is too slow there, so the CUDA path is checked against itself through
invariants the domain offers, plus a sampled comparison with the oracle.

* labels do not depend on how the batch is cut (partition invariance) nor on the precision path;
* a sample of the labels equals the oracle's;
* the flux matrix of unit weights is a transition COUNT matrix: integer entries, total = number of transitions,
  row sums = out-degree of every (overridden) parent label;
* accumulating iteration blocks one after the other gives bit for bit the matrix of one call over all of them
  (the serial association of the reference: get_fluxMatrix adds iteration matrices in order);
* weighted entries sum to the total weight (1e-12), and the sharded exchange formula (sum of block matrices)
  agrees to 1e-12;
* the page-locked-source path and the staged path of launch_ray_discretization give identical labels.
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg2():
    import torch

    import workloads as synthetic
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.engine import DeviceClusters

    cfg = synthetic.CONFIGS["cfg2"]
    dev = torch.device("cuda:0")
    means, centers = synthetic.make_centers(cfg)
    basis, target = synthetic.region_bounds(cfg)
    eng = DeviceClusters(RectilinearBinMapper(synthetic.boundaries(cfg)), centers, {b: b for b in range(cfg.n_bins)}, basis,
                         target, 1, device=dev)
    data = synthetic.generate_device(cfg, dev, means=means)
    return cfg, eng, data, centers


def test_labels_are_partition_and_path_invariant_and_match_oracle_sample(cfg2):
    import torch

    from msm_we_b200 import _lib

    cfg, eng, data, centers = cfg2
    X, pc = data["X"], data["pcoord"]
    full, bins, flags = eng.predict(X, pc, path=_lib.ASSIGN_FP64)
    tc, _, _ = eng.predict(X, pc, path=_lib.ASSIGN_TF32X3)
    assert torch.equal(full, tc)
    cut = 123457                                     # not a multiple of any tile size
    a, _, _ = eng.predict(X[:cut], pc[:cut], path=_lib.ASSIGN_FP64)
    b, _, _ = eng.predict(X[cut:], pc[cut:], path=_lib.ASSIGN_FP64)
    assert torch.equal(full, torch.cat([a, b]))
    eng.check_errors()
    # sampled oracle comparison (every 397th point)
    idx = torch.arange(0, X.shape[0], 397, device=X.device)
    Xs, bs, fs, ls = X[idx].cpu().numpy(), bins[idx].cpu().numpy(), flags[idx].cpu().numpy(), full[idx].cpu().numpy()
    T = sum(c.shape[0] for c in centers)
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in centers])])
    for i in range(len(ls)):
        if fs[i] & 2:
            assert ls[i] == T + 1
        elif fs[i] & 1:
            assert ls[i] == T
        else:
            lab, amb = O.kmeans_assign_tiebreak(Xs[i:i + 1], centers[bs[i]], return_ambiguous=True)
            assert amb[0] or ls[i] == offs[bs[i]] + lab[0]


def test_unit_weight_flux_is_an_exact_count_matrix(cfg2):
    import torch

    from msm_we_b200 import ops

    cfg, eng, data, _ = cfg2
    N = data["n"]
    labels, _, flags = eng.predict(data["X"], data["pcoord"])
    n = cfg.n_clusters
    ones = torch.ones(N, dtype=torch.float64, device=labels.device)
    dense = ops.flux_accumulate(labels[:N], labels[N:], ones, n, flag0=flags[:N], flag1=flags[N:],
                                   iter_offsets=data["iter_offsets"])
    d = dense.cpu().numpy()
    assert np.array_equal(d, np.rint(d)) and d.sum() == N
    # out-degree of every parent label after the overrides (basis parent -> n; a parent flag never becomes n+1)
    start = labels[:N].clone()
    start[(flags[:N] & 1) != 0] = n
    deg = np.bincount(start.cpu().numpy(), minlength=n + 2)
    assert np.array_equal(d.sum(axis=1), deg)


def test_flux_blockwise_accumulation_is_bit_identical_and_weights_are_conserved(cfg2):
    import torch

    from msm_we_b200 import ops

    cfg, eng, data, _ = cfg2
    N, w, offs = data["n"], data["weights"], data["iter_offsets"]
    labels, _, flags = eng.predict(data["X"], data["pcoord"])
    n = cfg.n_clusters
    start, end, f0, f1 = labels[:N], labels[N:], flags[:N], flags[N:]
    whole = ops.flux_accumulate(start, end, w, n, flag0=f0, flag1=f1, iter_offsets=offs)
    # same iterations in three consecutive blocks, accumulated into one running matrix
    running = torch.zeros_like(whole)
    separate = []
    I = offs.numel() - 1
    for lo, hi in ((0, 61), (61, 150), (150, I)):
        a, b = int(offs[lo]), int(offs[hi])
        sub = (offs[lo:hi + 1] - offs[lo]).contiguous()
        ops.flux_accumulate(start[a:b], end[a:b], w[a:b].contiguous(), n, flag0=f0[a:b], flag1=f1[a:b], iter_offsets=sub,
                            dense=running)
        blk = ops.flux_accumulate(start[a:b], end[a:b], w[a:b].contiguous(), n, flag0=f0[a:b], flag1=f1[a:b],
                                     iter_offsets=sub)
        separate.append(blk)
    assert torch.equal(whole, running)
    total = float(w.sum())
    assert abs(float(whole.sum()) - total) <= 1e-12 * total
    # the multi-GPU formula: block matrices summed in block order (what the peer-memory exchange computes)
    summed = (separate[0] + separate[1]) + separate[2]
    assert torch.allclose(summed, whole, rtol=1e-12, atol=1e-300)


def test_discretization_pinned_source_path_equals_staged_path(monkeypatch):
    import dataclasses

    import workloads as synthetic
    from msm_we_b200 import _pinning
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.msm_we import modelWE
    from msm_we_b200.stratified_clustering import StratifiedClusters

    cfg = dataclasses.replace(synthetic.CONFIGS["cfg2"], n_iters=40)
    means, centers = synthetic.make_centers(cfg)
    its = synthetic.generate_host(cfg, means)
    basis, target = synthetic.region_bounds(cfg)

    def run():
        model = modelWE()
        model.initialize(synthetic.to_iteration_source(its), None, "t", basis_pcoord_bounds=basis,
                         target_pcoord_bounds=target, tau=1.0, pcoord_ndim=1)
        model.get_iterations(); model.dimReduce()
        clusters = StratifiedClusters(RectilinearBinMapper(synthetic.boundaries(cfg)), model, cfg.k_per_bin, [])
        for b in range(cfg.n_bins):
            clusters.cluster_models[b].cluster_centers_ = centers[b]
        clusters.cluster_args["gpu_chunk_bytes"] = 8 << 20          # several chunks
        model.clusters = clusters; model.n_clusters = cfg.n_clusters
        model.launch_ray_discretization()
        model.get_fluxMatrix(n_lag=0, first_iter=0)
        return model

    monkeypatch.setattr(_pinning.PINS, "enabled", True)
    m1 = run()
    assert _pinning.PINS.bytes > 0                                   # the arrays really were page-locked in place
    monkeypatch.setattr(_pinning.PINS, "enabled", False)
    m2 = run()
    assert all(np.array_equal(a, b) for a, b in zip(m1.pair_dtrajs, m2.pair_dtrajs))
    assert np.array_equal(m1.fluxMatrixRaw, m2.fluxMatrixRaw)
