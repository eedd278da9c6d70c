"""GPU parity tests of the modelWE-level API (the drop-in boundary) against the CPU oracle.

Tolerances (from BASELINE.json north_star): cluster labels and transition counts bit-exact; centroids
and flux-matrix entries within 1e-12 relative (here they come out bit-exact on one GPU because the
kernels keep the serial reference's summation order).
"""
import copy
import pickle

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _build(cfg_name="tiny", use_weights=False):
    import workloads as synthetic
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.msm_we import modelWE

    cfg = synthetic.CONFIGS[cfg_name]
    means, centers = synthetic.make_centers(cfg)
    its = synthetic.generate_host(cfg, means)
    basis, target = synthetic.region_bounds(cfg)
    model = modelWE()
    model.initialize(synthetic.to_iteration_source(its), None, "synthetic", basis_pcoord_bounds=basis,
                     target_pcoord_bounds=target, tau=1.0, pcoord_ndim=1, use_weights_in_clustering=use_weights)
    model.get_iterations()
    model.dimReduce()
    mapper = RectilinearBinMapper(synthetic.boundaries(cfg))
    return cfg, model, mapper, its, centers, basis, target


class _OracleData:
    def __init__(self, its):
        self.its = its

    def iteration(self, n):
        it = self.its[n - 1]

        class R:
            pcoord0 = it["pcoord0"]; child = it["child"]; weights = it["weights"]
        return R


def _oracle_clustering(cfg, its, basis, target, use_weights, **cluster_args):
    """The reference flow with REAL sklearn models: batching rule of do_stratified_clustering, one
    MiniBatchKMeans.partial_fit per WE bin per batch (msm_we/_hamsm/_clustering.py:664-716, 890-916)."""
    from sklearn.cluster import MiniBatchKMeans

    om = O.RectilinearBinMapperOracle([np.append(np.arange(cfg.n_bins, dtype=np.float32), np.float32(np.inf))])
    args = {"n_clusters": cfg.k_per_bin, "max_iter": 100}
    args.update(cluster_args)
    models = [MiniBatchKMeans(**args) for _ in range(cfg.n_bins)]
    iters = list(range(1, cfg.n_iters))   # range(first_cluster_iter=1, maxIter)
    for used, batch in O.stratified_clustering_batches(_OracleData(its), iters, om, cfg.k_per_bin, basis, target,
                                                       use_weights=use_weights):
        for b, Xb, wb in batch:
            models[b].partial_fit(Xb, sample_weight=wb)
    return models, om


@pytest.mark.parametrize("use_weights", [False, True])
def test_cluster_coordinates_matches_sklearn_flow(use_weights):
    cfg, model, mapper, its, _, basis, target = _build("tiny", use_weights)
    model.cluster_coordinates(cfg.k_per_bin, stratified=True, use_ray=True, user_bin_mapper=mapper, random_state=1337)
    ref_models, om = _oracle_clustering(cfg, its, basis, target, use_weights, random_state=1337)
    assert model.n_clusters == cfg.k_per_bin * cfg.n_bins
    assert model.clustering_method == "stratified"
    n_fitted = 0
    for b in range(cfg.n_bins):
        ref_has = hasattr(ref_models[b], "cluster_centers_")
        assert hasattr(model.clusters.cluster_models[b], "cluster_centers_") == ref_has
        if ref_has:
            n_fitted += 1
            got = model.clusters.cluster_models[b]
            np.testing.assert_allclose(got.cluster_centers_, ref_models[b].cluster_centers_, rtol=RTOL, atol=1e-13)
            np.testing.assert_allclose(got._counts, ref_models[b]._counts, rtol=RTOL)
            assert got.n_steps_ == ref_models[b].n_steps_
            assert got._n_since_last_reassign == ref_models[b]._n_since_last_reassign
    assert n_fitted >= cfg.n_bins - 1
    # discretization with those centres: labels bit-exact against the oracle
    strat = O.StratifiedOracle(om, [getattr(m, "cluster_centers_", None) for m in ref_models], basis, target,
                               we_remap=model.clusters.we_remap)
    assert len(model.dtrajs) == cfg.n_iters - 1 and len(model.pair_dtrajs) == cfg.n_iters - 1
    for it in range(1, cfg.n_iters):
        d = its[it - 1]
        parent, child = O.discretize_iteration(strat, d["parent"], d["child"], d["pcoord0"], d["pcoord1"])
        assert np.array_equal(model.dtrajs[it - 1], child)
        assert np.array_equal(np.asarray(model.pair_dtrajs[it - 1]), np.stack([parent, child], axis=1))
    assert model.clusters.target_bins == strat.target_bins and model.clusters.basis_bins == strat.basis_bins


def _with_fixed_centers(cfg_name="tiny"):
    from msm_we_b200.stratified_clustering import StratifiedClusters

    cfg, model, mapper, its, centers, basis, target = _build(cfg_name)
    clusters = StratifiedClusters(mapper, model, cfg.k_per_bin, [])
    for b in range(cfg.n_bins):
        clusters.cluster_models[b].cluster_centers_ = centers[b]
    model.clusters = clusters
    model.n_clusters = cfg.k_per_bin * cfg.n_bins
    om = O.RectilinearBinMapperOracle(mapper.boundaries)
    return cfg, model, its, centers, basis, target, om


def test_predict_semantics_and_literal_reference_loop():
    cfg, model, its, centers, basis, target, om = _with_fixed_centers()
    clusters = model.clusters
    # make bin 5 unfitted and remapped to 4, as cluster_stratified does for never-filled bins
    del clusters.cluster_models[5].cluster_centers_
    clusters.we_remap[5] = 4
    cpb = [c if b != 5 else None for b, c in enumerate(centers)]
    strat = O.StratifiedOracle(om, cpb, basis, target, we_remap=clusters.we_remap)
    model.load_iter_data(3)
    d = its[2]
    clusters.processing_from = True
    got_parent = clusters.predict(d["parent"])
    clusters.processing_from = False
    got_child = clusters.predict(d["child"])
    # the literal reference loop: one sklearn predict([coord]) per segment
    ref_parent = strat.predict(d["parent"], d["pcoord0"], literal=True)
    ref_child = strat.predict(d["child"], d["pcoord1"], literal=True)
    assert got_parent.dtype == np.int64
    assert np.array_equal(got_parent, ref_parent) and np.array_equal(got_child, ref_child)
    T = sum(len(c) for c in cpb if c is not None)
    assert set(np.unique(got_child)) <= set(range(T + 2))
    # toggle alternates pcoord0List / pcoord1List (stratified_clustering.py:205-210)
    clusters.toggle = True
    clusters.processing_from = True
    a = clusters.predict(d["parent"])
    assert clusters.processing_from is False
    b = clusters.predict(d["child"])
    assert clusters.processing_from is True
    clusters.toggle = False
    assert np.array_equal(a, ref_parent) and np.array_equal(b, ref_child)
    # a free point in a bin without centres and without remap -> AssertionError (:187-189)
    clusters.we_remap[5] = 5
    model.load_iter_data(3)
    if np.any((om.assign(d["pcoord1"]) == 5)):
        clusters.processing_from = False
        with pytest.raises(AssertionError):
            clusters.predict(d["child"])


def test_parent_label_equals_previous_child_label():
    """Size-independent property of the synthetic data: a parent IS a child of the previous iteration."""
    cfg, model, its, centers, basis, target, om = _with_fixed_centers()
    model.launch_ray_discretization()
    for it in range(2, cfg.n_iters):
        prev_children = {tuple(np.round(x, 12)): l for x, l in zip(its[it - 2]["child"], model.dtrajs[it - 2])}
        pairs = np.asarray(model.pair_dtrajs[it - 1])
        for x, lab in zip(its[it - 1]["parent"][:40], pairs[:40, 0]):
            assert prev_children[tuple(np.round(x, 12))] == lab


def test_flux_matrix_bit_exact_and_iteration_conventions():
    cfg, model, its, centers, basis, target, om = _with_fixed_centers()
    model.launch_ray_discretization()
    model.get_fluxMatrix(n_lag=0)
    n = model.n_clusters
    # reference: iterations range(first_iter+1, maxIter) (_fluxmatrix.py:215); divide by their number
    iters = list(range(2, model.maxIter))
    per = [(np.asarray(model.pair_dtrajs[i - 1]), its[i - 1]["pcoord0"], its[i - 1]["pcoord1"], its[i - 1]["weights"])
           for i in iters]
    ref = O.flux_matrix(n, per, basis, target)
    assert model.fluxMatrixRaw.shape == (n + 2, n + 2)
    assert np.array_equal(model.fluxMatrixRaw, ref)
    assert model._fluxMatrixParams == [0, 1, None, None] and model.errorWeight == 0.0 and model.errorCount == 0
    # total weight is conserved: every iteration's weights sum to 1
    assert abs(model.fluxMatrixRaw.sum() - 1.0) < 1e-12
    # explicit iteration subsets, in chunks smaller than the data (chunked accumulation keeps the order)
    model.flux_chunk_transitions = 300
    model.get_fluxMatrix(n_lag=0, iters_to_use=[3, 5, 6])
    per = [per[i - 2] for i in (3, 5, 6)]
    assert np.array_equal(model.fluxMatrixRaw, O.flux_matrix(n, per, basis, target))
    one = model.get_iter_fluxMatrix(4)
    assert np.array_equal(one, O.iter_flux_matrix(n, np.asarray(model.pair_dtrajs[3]), its[3]["pcoord0"], its[3]["pcoord1"],
                                                  its[3]["weights"], basis, target))
    with pytest.raises(NotImplementedError):
        model.get_fluxMatrix(n_lag=1)


def test_build_flux_matrix_static_matches_scipy_reference():
    from msm_we_b200.msm_we import modelWE

    rng = np.random.default_rng(3)
    n, S = 50, 400
    pairs = rng.integers(0, n, size=(S, 2))
    w = rng.uniform(size=S)
    sb = np.where(rng.uniform(size=S) < 0.1); eb = np.where(rng.uniform(size=S) < 0.1); et = np.where(rng.uniform(size=S) < 0.1)
    got = modelWE.build_flux_matrix(n, pairs, sb, eb, et, w)
    ref = O.build_flux_matrix(n, pairs, sb, eb, et, w)
    assert got.shape == ref.shape == (n + 2, n + 2)
    assert np.array_equal(np.asarray(got.todense()), np.asarray(ref.todense()))
    got2, it = modelWE.build_flux_matrix_remote.remote(n, pairs, sb, eb, et, w, 7)
    assert it == 7 and np.array_equal(np.asarray(got2.todense()), np.asarray(ref.todense()))
    bad = pairs.copy(); bad[0, 0] = n + 5
    with pytest.raises(ValueError):
        modelWE.build_flux_matrix(n, bad, (np.array([], dtype=int),), (np.array([], dtype=int),), (np.array([], dtype=int),), w)


def test_model_pickles_and_deepcopies_without_gpu_state():
    cfg, model, its, centers, basis, target, om = _with_fixed_centers()
    model.launch_ray_discretization()
    assert model.clusters._device is not None
    clone = pickle.loads(pickle.dumps(model))
    assert clone.clusters._device is None
    deep = copy.deepcopy(model)
    deep.load_iter_data(2)
    deep.clusters.model = deep
    assert np.array_equal(deep.clusters.predict(its[1]["child"]), model.dtrajs[1])
    # reassigning centres (organize_stratified does np.delete + assignment) invalidates the device snapshot
    cm = model.clusters.cluster_models[2]
    cm.cluster_centers_ = np.delete(cm.cluster_centers_, [0, 3], 0)
    model.load_iter_data(2)
    lab = model.clusters.predict(its[1]["child"])
    cpb = [c for c in model.clusters.centers_per_bin()]
    strat = O.StratifiedOracle(om, cpb, basis, target)
    assert np.array_equal(lab, strat.predict(its[1]["child"], its[1]["pcoord1"]))


def test_do_stratified_ray_discretization_remote_shim():
    cfg, model, its, centers, basis, target, om = _with_fixed_centers()
    res = model.do_stratified_ray_discretization.remote(model, model.clusters, 4, model.processCoordinates)
    (parent, child), one, it, tb, bb = res
    strat = O.StratifiedOracle(om, centers, basis, target)
    rp, rc = O.discretize_iteration(strat, its[3]["parent"], its[3]["child"], its[3]["pcoord0"], its[3]["pcoord1"])
    assert one == 1 and it == 4 and np.array_equal(parent, rp) and np.array_equal(child, rc)
    assert tb == strat.target_bins and bb == strat.basis_bins


def test_ntl9_fixture_replay_sparsity(golden_dir):
    """BASELINE config 1 substitute: the reference's pickled pair_dtrajs replayed through K3 reproduce the
    non-zero pattern of its golden fluxmatrix_raw.npy (weights are not in the pickle)."""
    import torch
    from msm_we_b200 import ops

    g = np.load(f"{golden_dir}/ntl9_clustered.npz")
    n = int(g["n_clusters"])
    lens = g["pair_lens"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    # get_fluxMatrix skips iteration 1 (range(first_iter+1, maxIter)); predict labelled basis/target T, T+1 = 275, 276
    s = g["pair_parent"][offs[1]:].copy(); e = g["pair_child"][offs[1]:].copy()
    T = 275
    for arr in (s, e):
        arr[arr == T] = n
        arr[arr == T + 1] = n + 1
    dev = torch.device("cuda:0")
    dense = ops.flux_accumulate(torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev), None, n).cpu().numpy()
    got = set(zip(*np.nonzero(dense)))
    ref = set(zip(g["flux_raw_nz_i"].tolist(), g["flux_raw_nz_j"].tolist()))
    assert got == ref and len(ref) == 4575
    assert dense.sum() == len(s) == 10340


def test_hotpath_step_matches_separate_kernels_and_oracle():
    """The one-call C entry (K0 -> K1 -> K3 -> / nI) on a stacked batch equals the modelWE-level results."""
    import torch
    import workloads as synthetic

    cfg, model, its, centers, basis, target, om = _with_fixed_centers()
    model.launch_ray_discretization()
    model.get_fluxMatrix(n_lag=0, first_iter=0)          # iterations 1 .. maxIter-1, as the stacked batch below
    dev = model.clusters.device_state()
    use = its[: cfg.n_iters - 1]
    X2 = torch.from_numpy(np.concatenate([d["parent"] for d in use] + [d["child"] for d in use])).to(dev.device)
    P2 = torch.from_numpy(np.concatenate([d["pcoord0"] for d in use] + [d["pcoord1"] for d in use])).to(dev.device)
    w = torch.from_numpy(np.concatenate([d["weights"] for d in use])).to(dev.device)
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum([len(d["weights"]) for d in use])]).astype(np.int64)).to(dev.device)
    M = model.n_clusters + 2
    dense = torch.zeros((M, M), dtype=torch.float64, device=dev.device)
    labels = dev.hotpath_step(X2, P2, w, model.n_clusters, iter_offsets=offs, dense=dense, divisor=float(len(use)))
    dev.check_errors()
    n = w.numel()
    lab = labels.cpu().numpy()
    assert np.array_equal(lab[n:], np.concatenate(model.dtrajs))
    assert np.array_equal(lab[:n], np.concatenate([np.asarray(p)[:, 0] for p in model.pair_dtrajs]))
    assert np.array_equal(dense.cpu().numpy(), model.fluxMatrixRaw)


@pytest.mark.gpu
def test_peer_memory_flux_allreduce_two_gpus():
    """The one-kernel exchange step (csrc/peer_reduce.cu) on 2 GPUs of one node: every rank ends with exactly the
    rank-order sum divided by nI (numpy on the gathered partials), and agrees with NCCL to rounding.  Skipped on
    single-GPU boxes; `gpurun --gpus 2 -- python -m pytest tests -m gpu -k peer` runs it."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "peer_reduce_check.py")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "peer path available: True" in out.stdout
    assert "values identical on all ranks: True" in out.stdout


@pytest.mark.gpu
def test_discretization_with_device_projection_matches_host_projection():
    """A fitted linear projection (the reference's PCA coordinates.transform) in front of predict: shipping the raw
    features and projecting on the device labels every frame like projecting on the host first."""
    from msm_we_b200.msm_we import LinearCoordinates
    from msm_we_b200.stratified_clustering import StratifiedClusters

    cfg, model, mapper, its, centers, basis, target = _build("tiny")
    rng = np.random.default_rng(3)
    d_out = 6
    comps = np.linalg.qr(rng.normal(size=(cfg.dim, cfg.dim)))[0][:d_out]
    mean = rng.normal(size=cfg.dim)
    proj_centers = [(c - mean) @ comps.T for c in centers]

    class HostOnly:                       # same transform, no device_projection attribute -> host matmul path
        def __init__(self, lc):
            self.lc = lc

        def transform(self, x):
            return self.lc.transform(x)

    labels = []
    for coords in (LinearCoordinates(comps, mean), HostOnly(LinearCoordinates(comps, mean))):
        model.coordinates = coords
        clusters = StratifiedClusters(mapper, model, cfg.k_per_bin, [])
        for b in range(cfg.n_bins):
            clusters.cluster_models[b].cluster_centers_ = proj_centers[b]
        model.clusters = clusters
        model.n_clusters = cfg.n_clusters
        model.launch_ray_discretization()
        labels.append(np.concatenate(model.pair_dtrajs))
    same = labels[0] == labels[1]
    assert same.mean() > 0.9999              # projections differ by rounding only: at most a near-tie may flip
    # and both agree with the oracle on the host-projected coordinates wherever the oracle is unambiguous
    it = its[1]
    Xp = O.linear_transform(it["child"], comps, mean)
    ref = O.StratifiedOracle(O.RectilinearBinMapperOracle([np.append(np.arange(cfg.n_bins, dtype=np.float32), np.float32(np.inf))]),
                             proj_centers, basis, target).predict(Xp, it["pcoord1"])
    assert (model.pair_dtrajs[1][:, 1] == ref).mean() > 0.999
