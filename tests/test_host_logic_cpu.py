"""CPU: host-side logic of the drop-in layer (no kernels): bin mappers, the MiniBatchKMeans bookkeeping and
RNG stream of BinClusterModel, the batching rule of do_stratified_clustering, iteration sharding and the
all-reduce orchestration (gloo, world_size 2)."""
import os
import pickle
import sys

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bin_mappers_match_oracle():
    from msm_we_b200.binning import RectilinearBinMapper, VoronoiBinMapper, mapper_kind

    rng = np.random.default_rng(0)
    b = [np.array([0, .2, .25, .3, np.inf]), np.linspace(-1, 1, 5)]
    m = RectilinearBinMapper(b)
    pc = np.stack([rng.uniform(0, 5, 500), rng.uniform(-1, 0.999, 500)], axis=1)
    assert m.nbins == 16 and np.array_equal(m.assign(pc), O.RectilinearBinMapperOracle(b).assign(pc))
    with pytest.raises(ValueError):
        m.assign(np.array([[0.1, 1.0]]))
    assert mapper_kind(m) == "rectilinear"
    c = rng.normal(size=(9, 2))
    v = VoronoiBinMapper(centers=c)
    assert np.array_equal(v.assign(pc), O.VoronoiBinMapperOracle(c).assign(pc)) and mapper_kind(v) == "voronoi"
    assert mapper_kind(VoronoiBinMapper(dfunc=lambda p, cs: np.abs(cs - p).sum(axis=1), centers=c)) == "host"


def _patch_partial_fit_with_oracle(monkeypatch):
    """Run the host flow with the oracle's numpy E/M steps standing in for the K1/K2 kernels."""
    import msm_we_b200.clustering_ops as co

    def fake(batch, device=None):
        for model, X, w in batch:
            Xc, wc, reassign = model._prepare(X, w)
            labels = O.kmeans_assign_tiebreak(Xc, model.cluster_centers_)
            O.minibatch_update(Xc, wc, model.cluster_centers_, model._counts, labels)
            model._finish(Xc, reassign)

    monkeypatch.setattr(co, "partial_fit_models", fake)


@pytest.mark.parametrize("init", ["k-means++", "random"])
@pytest.mark.parametrize("weighted", [False, True])
def test_bincluster_model_follows_sklearn_rng_and_bookkeeping(monkeypatch, init, weighted):
    from sklearn.cluster import MiniBatchKMeans
    from msm_we_b200.stratified_clustering import BinClusterModel

    _patch_partial_fit_with_oracle(monkeypatch)
    rng = np.random.default_rng(3)
    K, D = 10, 6
    ours = BinClusterModel(n_clusters=K, init=init, random_state=42, max_iter=100)
    ref = MiniBatchKMeans(n_clusters=K, init=init, random_state=42, max_iter=100)
    for step in range(6):
        n = int(rng.integers(K, 80))
        X = rng.normal(size=(n, D)) + 2 * rng.integers(0, 3, size=(n, 1))
        w = np.exp(rng.normal(0, 2, size=n)) if weighted else None
        ours.partial_fit(X, sample_weight=w)
        ref.partial_fit(X, sample_weight=w)
        assert np.array_equal(ours.cluster_centers_, ref.cluster_centers_), (step,)
        assert np.array_equal(ours._counts, ref._counts)
        assert ours._n_since_last_reassign == ref._n_since_last_reassign and ours.n_steps_ == ref.n_steps_
        assert ours._batch_size == ref._batch_size and ours._init_size == ref._init_size
    with pytest.raises(ValueError):
        BinClusterModel(n_clusters=50).partial_fit(np.zeros((3, 2)))


def test_cluster_stratified_host_flow_matches_sklearn_flow(monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_model_gpu as T

    _patch_partial_fit_with_oracle(monkeypatch)
    for use_weights in (False, True):
        cfg, model, mapper, its, _, basis, target = T._build("tiny", use_weights)
        monkeypatch.setattr(type(model), "launch_ray_discretization", lambda self, progress_bar=None: None)
        model.cluster_coordinates(cfg.k_per_bin, stratified=True, use_ray=True, user_bin_mapper=mapper, random_state=1337)
        ref_models, _ = T._oracle_clustering(cfg, its, basis, target, use_weights, random_state=1337)
        for b in range(cfg.n_bins):
            assert hasattr(model.clusters.cluster_models[b], "cluster_centers_") == hasattr(ref_models[b], "cluster_centers_")
            if hasattr(ref_models[b], "cluster_centers_"):
                assert np.array_equal(model.clusters.cluster_models[b].cluster_centers_, ref_models[b].cluster_centers_)
        assert model.n_clusters == cfg.k_per_bin * cfg.n_bins
        state = pickle.loads(pickle.dumps(model.clusters))
        assert state._device is None and len(state.cluster_models) == cfg.n_bins


def test_find_nearest_bin_matches_oracle():
    from msm_we_b200._hamsm._clustering import ClusteringMixin
    from msm_we_b200.binning import RectilinearBinMapper

    m = RectilinearBinMapper([np.array([0, 1, 2, 3, 4, 5.0])])
    om = O.RectilinearBinMapperOracle([np.array([0, 1, 2, 3, 4, 5.0])])
    for filled in ([0, 4], [1, 2], [3]):
        for b in range(5):
            if b not in filled:
                assert ClusteringMixin.find_nearest_bin(m, b, filled) == O.find_nearest_bin(om, b, filled)
    with pytest.raises(AssertionError):
        ClusteringMixin.find_nearest_bin(m, 0, [])


def test_partition_iterations_is_contiguous_and_balanced():
    from msm_we_b200.distributed import partition_iterations

    iters = list(range(2, 100))
    counts = np.linspace(10, 110, len(iters))          # segments grow during a WE run
    for world in (1, 2, 4, 8):
        parts = partition_iterations(iters, counts, world)
        assert sum(parts, []) == iters
        loads = [sum(counts[i - 2] for i in p) for p in parts]
        assert max(loads) - min(loads) <= 2 * counts.max()
    assert partition_iterations([5], [3], 4)[0] + sum(partition_iterations([5], [3], 4)[1:], []) == [5]


def _gloo_worker(rank, world, port, tmp):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import test_model_gpu as T
    from msm_we_b200.distributed import get_fluxMatrix_sharded

    cfg, model, mapper, its, centers, basis, target = T._build("tiny")
    om = O.RectilinearBinMapperOracle(mapper.boundaries)
    strat = O.StratifiedOracle(om, centers, basis, target)
    model.n_clusters = cfg.n_clusters
    model.pair_dtrajs = [np.stack(O.discretize_iteration(strat, d["parent"], d["child"], d["pcoord0"], d["pcoord1"]), axis=1)
                         for d in its[: cfg.n_iters - 1]]

    def local_flux(m, iters):          # the oracle stands in for K0 + K3 on this rank's iteration block
        M = m.n_clusters + 2
        tot = np.zeros((M, M))
        for i in iters:
            d = its[i - 1]
            tot = tot + O.iter_flux_matrix(m.n_clusters, m.pair_dtrajs[i - 1], d["pcoord0"], d["pcoord1"], d["weights"], basis, target)
        return torch.from_numpy(tot)

    out = get_fluxMatrix_sharded(model, n_lag=0, local_flux_fn=local_flux)
    np.save(os.path.join(tmp, f"flux_{rank}.npy"), out)
    dist.destroy_process_group()


def test_sharded_flux_matrix_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_model_gpu as T

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "flux_0.npy"), np.load(tmp_path / "flux_1.npy")
    assert np.array_equal(a, b)
    cfg, model, mapper, its, centers, basis, target = T._build("tiny")
    om = O.RectilinearBinMapperOracle(mapper.boundaries)
    strat = O.StratifiedOracle(om, centers, basis, target)
    per = []
    for i in range(2, cfg.n_iters):
        d = its[i - 1]
        pr, ch = O.discretize_iteration(strat, d["parent"], d["child"], d["pcoord0"], d["pcoord1"])
        per.append((np.stack([pr, ch], axis=1), d["pcoord0"], d["pcoord1"], d["weights"]))
    ref = O.flux_matrix(cfg.n_clusters, per, basis, target)
    # two partial sums instead of one serial sum: 1e-12 relative is the stated multi-GPU tolerance
    assert np.allclose(a, ref, rtol=1e-12, atol=0) and abs(a.sum() - 1.0) < 1e-12


def test_lean_flux_gather_equals_full_gather_and_keeps_seg_weights():
    """get_fluxMatrix loads all but the last iteration of a pass through a lean gather: same four arrays as the
    reference-shaped loader, NaN-coordinate segments zero-weighted, seg_weights recorded."""
    import dataclasses

    import workloads as synthetic
    from msm_we_b200.msm_we import modelWE

    cfg = dataclasses.replace(synthetic.CONFIGS["tiny"], n_iters=6)
    means, _ = synthetic.make_centers(cfg)
    its = synthetic.generate_host(cfg, means)
    its[2]["child"][5, 3] = np.nan                      # a broken frame in iteration 3
    basis, target = synthetic.region_bounds(cfg)
    model = modelWE()
    model.initialize(synthetic.to_iteration_source(its), None, "t", basis_pcoord_bounds=basis, target_pcoord_bounds=target,
                     tau=1.0, pcoord_ndim=1)
    model.get_iterations()
    rng = np.random.default_rng(0)
    model.pair_dtrajs = [rng.integers(0, 10, size=(cfg.n_segs, 2)) for _ in range(cfg.n_iters)]
    for n_iter in (1, 3, 5):
        full = model._gather_flux_inputs(n_iter)
        model.seg_weights.pop(n_iter)
        lean = model._gather_flux_inputs_lean(n_iter)
        for a, b in zip(full, lean):
            assert np.array_equal(np.asarray(a), np.asarray(b))
        assert np.array_equal(model.seg_weights[n_iter], its[n_iter - 1]["weights"])
    assert model._gather_flux_inputs_lean(3)[3][5] == 0.0


def test_host_pinning_degrades_without_a_gpu():
    """HostPins.ensure never raises: arrays it cannot page-lock (no device here, wrong dtype/layout, disabled) are
    reported as not pinned and the caller stages them."""
    from msm_we_b200._pinning import HostPins

    pins = HostPins()
    a = np.zeros((64, 8))
    assert pins.ensure(a[:, :4]) is False               # not contiguous
    assert pins.ensure(a.astype(np.float32)) is False   # not float64
    assert pins.ensure("nope") is False
    got = pins.ensure(a)                                # needs a CUDA device: False here, never an exception
    assert got in (False, True)
    pins.enabled = False
    assert pins.ensure(np.ones(16)) is False
    view = a.reshape(-1)[8:24]
    assert HostPins._owner(view) is a


def _gloo_lloyd_worker(rank, world, port, tmp, mode="single"):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpu_emulation import emulate_kernels
    from msm_we_b200 import clustering_ops

    mp_ = pytest.MonkeyPatch()
    emulate_kernels(mp_)
    if mode == "multi_ranked":
        mp_.setattr(clustering_ops, "EXACT_ORDER_MAX_POINTS", 0)  # the per-bin top-k path instead of numpy's argpartition
    X, bins, init, K = _lloyd_case(multi=mode != "single")
    n = len(X)
    lo, hi = n * rank // world, n * (rank + 1) // world          # contiguous shard, as iteration ranges are
    centers = torch.from_numpy(np.concatenate(init).copy())
    offs = torch.from_numpy(np.arange(0, (len(init) + 1) * K, K, dtype=np.int64))
    clustering_ops.lloyd_fit(torch.from_numpy(X[lo:hi].copy()), None, torch.from_numpy(bins[lo:hi].copy()), centers, offs, K, 4,
                             group=dist.group.WORLD)
    np.save(os.path.join(tmp, f"centers_{rank}.npy"), centers.numpy())
    mp_.undo()
    dist.destroy_process_group()


def _lloyd_case(multi=False):
    rng = np.random.default_rng(21)
    nbins, K, D = 3, 5, 6
    Xs, bs, init = [], [], []
    for b in range(nbins):
        m = rng.normal(0, 4, size=(K, D))
        k = rng.integers(0, K, size=300)
        Xs.append(m[k] + rng.normal(0, 0.7, size=(300, D)))
        bs.append(np.full(300, b, dtype=np.int32))
        c = m + rng.normal(0, 0.3, size=(K, D))
        if b == 1:
            c[3] = c[0]              # a duplicate centre: empty cluster -> relocation, candidates exchanged between ranks
            if multi:
                c[2] = c[0]          # two empty clusters in one model: the bin's two farthest points, over both ranks
        init.append(c)
    order = rng.permutation(900)
    return np.concatenate(Xs)[order], np.concatenate(bs)[order], init, K


@pytest.mark.parametrize("mode", ["single", "multi", "multi_ranked"])
def test_sharded_lloyd_gloo_world2_matches_single_process(tmp_path, monkeypatch, mode):
    """Iteration-range sharded Lloyd (partial sums all-reduced, empty-cluster candidates all-gathered) equals the
    single-process fit and the real sklearn KMeans per bin.  With several empty clusters in one model the relocated
    centres may land on permuted cluster indices of that bin (clustering_ops._relocate_empty_clusters): compared as
    sets of rows."""
    import torch
    import torch.multiprocessing as mp
    from sklearn.cluster import KMeans

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    port = 29500 + ((os.getpid() + 977) % 2000)
    mp.spawn(_gloo_lloyd_worker, args=(2, port, str(tmp_path), mode), nprocs=2, join=True)
    a, b = np.load(tmp_path / "centers_0.npy"), np.load(tmp_path / "centers_1.npy")
    assert np.array_equal(a, b)
    canon = (lambda m: m) if mode == "single" else (lambda m: m[np.lexsort(m.T[::-1])])
    from cpu_emulation import emulate_kernels
    from msm_we_b200 import clustering_ops

    emulate_kernels(monkeypatch)
    X, bins, init, K = _lloyd_case(multi=mode != "single")
    centers = torch.from_numpy(np.concatenate(init).copy())
    offs = torch.from_numpy(np.arange(0, (len(init) + 1) * K, K, dtype=np.int64))
    clustering_ops.lloyd_fit(torch.from_numpy(X), None, torch.from_numpy(bins), centers, offs, K, 4)
    single = centers.numpy()
    for bb in range(len(init)):
        sl = slice(bb * K, (bb + 1) * K)
        assert np.allclose(canon(a[sl]), canon(single[sl]), rtol=1e-12, atol=1e-13), bb
        ref = KMeans(n_clusters=K, init=init[bb], n_init=1, max_iter=4, tol=0.0, algorithm="lloyd").fit(X[bins == bb]).cluster_centers_
        assert np.allclose(canon(a[sl]), canon(ref), rtol=1e-11, atol=1e-12), bb
