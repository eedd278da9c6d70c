"""Full Lloyd iterations (SURVEY section 8 row a5''; BASELINE config 5's "10 Lloyd iters"): ``clustering_ops.lloyd_fit`` and
``modelWE.lloyd_refine_clusters`` against the REAL scikit-learn ``KMeans`` (init = fixed centres, n_init = 1,
algorithm = "lloyd", tol = 0), which is what the reference calls (msm_we/_hamsm/_clustering.py:289,491), one model per
WE bin.  Centroids to 1e-12 relative (sklearn centres X before fitting, so the last bits differ), final labels exact.
Runs through the CUDA library with ``-m gpu`` and through the numpy stand-ins (host logic) otherwise."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

BACKENDS = [pytest.param(False, id="host-logic-cpu"), pytest.param(True, id="cuda", marks=pytest.mark.gpu)]


def _backend(monkeypatch, gpu):
    import torch

    if gpu:
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        return torch.device("cuda", 0)
    from cpu_emulation import emulate_kernels

    emulate_kernels(monkeypatch)
    return torch.device("cpu")


def _sklearn_lloyd(X, init, n_iter, w=None):
    from sklearn.cluster import KMeans

    km = KMeans(n_clusters=init.shape[0], init=init, n_init=1, max_iter=n_iter, tol=0.0, algorithm="lloyd")
    km.fit(X, sample_weight=w)
    return km.cluster_centers_, km.predict(X)


def _case(rng, nbins, K, D, per_bin, spread=2.0):
    means = rng.normal(0, 3, size=(nbins, 1, D)) + rng.normal(0, spread, size=(nbins, K, D))
    Xs, bins = [], []
    for b in range(nbins):
        n = int(per_bin * rng.uniform(0.5, 1.5))
        k = rng.integers(0, K, size=n)
        Xs.append(means[b, k] + rng.normal(0, 1.0, size=(n, D)))
        bins.append(np.full(n, b, dtype=np.int32))
    X = np.concatenate(Xs)
    bins = np.concatenate(bins)
    order = rng.permutation(len(bins))
    init = [means[b] + rng.normal(0, 0.5, size=(K, D)) for b in range(nbins)]
    return X[order], bins[order], init


@pytest.mark.parametrize("gpu", BACKENDS)
@pytest.mark.parametrize("nbins,K,D,per_bin,weighted", [(5, 8, 16, 400, False), (3, 20, 64, 900, True), (4, 100, 256, 1500, False)])
def test_lloyd_fit_ten_iterations_match_sklearn_kmeans(monkeypatch, gpu, nbins, K, D, per_bin, weighted):
    import torch

    dev = _backend(monkeypatch, gpu)
    if not gpu and K * D > 2000:
        pytest.skip("the pure-Python stand-in is too slow for the config-5 shape; covered by the cuda variant")
    from msm_we_b200 import clustering_ops, ops

    rng = np.random.default_rng(nbins * 100 + K)
    X, bins, init = _case(rng, nbins, K, D, per_bin)
    w = np.exp(rng.normal(0, 1, size=len(X))) if weighted else None
    centers = torch.from_numpy(np.concatenate(init)).to(dev)
    offs = torch.from_numpy(np.arange(0, (nbins + 1) * K, K, dtype=np.int64)).to(dev)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    clustering_ops.lloyd_fit(t(X), t(w), t(bins), centers, offs, K, 10)
    labels = ops.assign_stratified(t(X), t(bins), None, centers, ops.centers_sqnorm(centers), offs, K).cpu().numpy()
    got = centers.cpu().numpy()
    for b in range(nbins):
        sel = bins == b
        ref_c, ref_l = _sklearn_lloyd(X[sel], init[b], 10, None if w is None else w[sel])
        assert np.allclose(got[b * K:(b + 1) * K], ref_c, rtol=1e-12, atol=1e-13), f"bin {b}: centroids differ"
        assert np.array_equal(labels[sel] - b * K, ref_l), f"bin {b}: final labels differ"


@pytest.mark.parametrize("gpu", BACKENDS)
def test_lloyd_relocates_empty_clusters_like_sklearn(monkeypatch, gpu):
    """Duplicate initial centres: ties go to the lower index, the duplicate receives nothing and sklearn moves it onto
    the point farthest from its own centre (_relocate_empty_clusters_dense)."""
    import torch

    dev = _backend(monkeypatch, gpu)
    from msm_we_b200 import clustering_ops

    rng = np.random.default_rng(5)
    K, D = 6, 8
    X = np.concatenate([rng.normal(m, 0.5, size=(60, D)) for m in (-4, 0, 4, 9)])
    init = np.stack([X[0], X[0], X[70], X[70], X[130], X[200]])        # two duplicated pairs -> two empty clusters
    for n_iter in (1, 4):
        centers = torch.from_numpy(init.copy()).to(dev)
        offs = torch.tensor([0, K], dtype=torch.int64, device=dev)
        bins = torch.zeros(len(X), dtype=torch.int32, device=dev)
        clustering_ops.lloyd_fit(torch.from_numpy(X).to(dev), None, bins, centers, offs, K, n_iter)
        ref_c, _ = _sklearn_lloyd(X, init, n_iter)
        assert np.allclose(centers.cpu().numpy(), ref_c, rtol=1e-11, atol=1e-12), n_iter


@pytest.mark.parametrize("gpu", BACKENDS)
def test_model_lloyd_refine_clusters_matches_sklearn_per_bin(monkeypatch, gpu):
    """Through the modelWE API on host arrays: frames binned by the PARENT pcoord, basis / target parents left out."""
    _backend(monkeypatch, gpu)
    import test_model_gpu as T
    from oracle import oracle as O

    cfg, model, mapper, its, centers, basis, target = T._build("tiny")
    from msm_we_b200.stratified_clustering import StratifiedClusters

    clusters = StratifiedClusters(mapper, model, cfg.k_per_bin, [])
    for b in range(cfg.n_bins):
        clusters.cluster_models[b].cluster_centers_ = centers[b].copy()
    model.clusters = clusters
    model.n_clusters = cfg.n_clusters
    used = model.lloyd_refine_clusters(3)
    X = np.concatenate([d["child"] for d in its[: cfg.n_iters - 1]])
    pc0 = np.concatenate([d["pcoord0"] for d in its[: cfg.n_iters - 1]])
    assert used == len(X)
    keep = ~(O.is_we_region(pc0, basis) | O.is_we_region(pc0, target))
    bins = O.RectilinearBinMapperOracle(mapper.boundaries).assign(pc0)
    for b in range(cfg.n_bins):
        rows = X[keep & (bins == b)]
        if len(rows) < cfg.k_per_bin:
            continue
        ref_c, _ = _sklearn_lloyd(rows, centers[b], 3)
        assert np.allclose(clusters.cluster_models[b].cluster_centers_, ref_c, rtol=1e-11, atol=1e-12), b


@pytest.mark.parametrize("gpu", BACKENDS)
def test_discretization_after_lloyd_reuses_the_resident_child_rows(monkeypatch, gpu):
    """lloyd_refine_clusters leaves the end-of-segment feature rows on the device for the discretization that follows
    (single use); labels must equal a pass that ships everything itself, and a copy of the model takes no device rows."""
    import copy, pickle

    _backend(monkeypatch, gpu)
    import test_model_gpu as T
    from msm_we_b200.stratified_clustering import StratifiedClusters

    def prepared(refine, **cluster_args):
        cfg, model, mapper, its, centers, basis, target = T._build("tiny")
        clusters = StratifiedClusters(mapper, model, cfg.k_per_bin, [])
        clusters.cluster_args.update(cluster_args)
        for b in range(cfg.n_bins):
            clusters.cluster_models[b].cluster_centers_ = centers[b].copy()
        model.clusters = clusters
        model.n_clusters = cfg.n_clusters
        if refine:
            model.lloyd_refine_clusters(2, iters_to_use=refine)
        return cfg, model

    cfg, model = prepared(None)
    every = list(range(1, model.maxIter))
    some = every[1::2]                                   # only part of the iterations resident
    results = []
    for refine, args in ((every, {}), (some, {}), (every, {"gpu_resident_bytes": 0})):
        cfg, m = prepared(refine, **args)
        if args:
            assert m._resident_child_rows is None        # over budget: nothing kept
        else:
            assert m._resident_child_rows is not None and set(m._resident_child_rows.rows) == set(refine)
            assert copy.deepcopy(m._resident_child_rows) is None and pickle.loads(pickle.dumps(m._resident_child_rows)) is None
        centres = [c.cluster_centers_.copy() for c in m.clusters.cluster_models]
        m.launch_ray_discretization()
        assert m._resident_child_rows is None            # consumed
        results.append((refine, centres, [d.copy() for d in m.pair_dtrajs]))
    # reference for each: the same refined centres, every frame shipped by the discretization itself
    for refine, centres, pairs in results:
        cfg, m = prepared(None)
        for c, v in zip(m.clusters.cluster_models, centres):
            c.cluster_centers_ = v
        m.launch_ray_discretization()
        assert len(pairs) == len(m.pair_dtrajs)
        for a, b in zip(pairs, m.pair_dtrajs):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("dups", [2, 9])          # 3 empty clusters: per-bin top-k kernel; 10: the two-sort fallback (k > 8)
@pytest.mark.parametrize("gpu", BACKENDS)
def test_lloyd_relocation_device_ranked_branch_relocates_the_same_points(monkeypatch, gpu, dups):
    """Large models rank the farthest points on the device (clustering_ops.EXACT_ORDER_MAX_POINTS): the same points are
    relocated as by sklearn; only the cluster index that receives each of them may be permuted within the bin."""
    import torch

    dev = _backend(monkeypatch, gpu)
    from msm_we_b200 import clustering_ops

    monkeypatch.setattr(clustering_ops, "EXACT_ORDER_MAX_POINTS", 0)
    rng = np.random.default_rng(8)
    K, D = dups + 5, 5
    X = np.concatenate([rng.normal(m, 0.4, size=(80, D)) for m in (-5, 0, 5, 11)])
    init = np.stack([X[0]] * (dups + 1) + [X[90], X[90], X[170], X[250]])  # dups + 1 empty clusters in one model
    centers = torch.from_numpy(init.copy()).to(dev)
    offs = torch.tensor([0, K], dtype=torch.int64, device=dev)
    bins = torch.zeros(len(X), dtype=torch.int32, device=dev)
    clustering_ops.lloyd_fit(torch.from_numpy(X).to(dev), None, bins, centers, offs, K, 1)
    ref_c, _ = _sklearn_lloyd(X, init, 1)
    got = centers.cpu().numpy()
    key = lambda a: a[np.lexsort(a.T[::-1])]            # noqa: E731  rows in a canonical order
    assert np.allclose(key(got), key(ref_c), rtol=1e-11, atol=1e-12)
