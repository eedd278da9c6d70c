"""Parity against fixtures produced by EXECUTING the reference's own code (tests/golden/make_reference_fixtures.py,
tests/golden/refshim.py): the same WE data goes through this package's modelWE API and every result the reference
left on its model is compared -- labels, transition pairs and integer bookkeeping bit-exact; centroids, counts and
flux-matrix entries to 1e-12 relative (the north star's tolerance) with identical sparsity patterns.

Each check exists twice: ``-m gpu`` runs it through the CUDA library (the parity test proper); ``-m "not gpu"`` runs
the same host code with the numpy stand-ins of tests/cpu_emulation.py in place of the kernels, which pins the host
logic (HDF5 feeder, batch planning, staging, cleaning, block validation) in the CPU-only container.
"""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
for p in (HERE, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)

import fixture_data as FD  # noqa: E402
import refshim  # noqa: E402

RTOL = 1e-12


def _backend(request, monkeypatch, gpu):
    if gpu:
        import torch

        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
    else:
        from cpu_emulation import emulate_kernels

        emulate_kernels(monkeypatch)
    # the product's HDF5 feeder imports h5py; the fixtures' WE files live in refshim's in-memory registry
    monkeypatch.setitem(sys.modules, "h5py", refshim.fake_h5py_module())


BACKENDS = [pytest.param(False, id="host-logic-cpu"), pytest.param(True, id="cuda", marks=pytest.mark.gpu)]


def _close(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
    assert np.array_equal(a == 0, b == 0), f"{what}: zero pattern differs"
    ok = np.isclose(a, b, rtol=RTOL, atol=0, equal_nan=True)
    assert ok.all(), f"{what}: max rel diff {np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)):.3e}"


def _mapper(fx):
    from msm_we_b200.binning import RectilinearBinMapper, VoronoiBinMapper

    if "voronoi_centers" in fx.files:
        return VoronoiBinMapper(centers=fx["voronoi_centers"])     # (the default distance: Euclidean, evaluated by K0)
    bnds, p = [], 0
    for ln in fx["boundary_lens"]:
        bnds.append(fx["boundaries"][p:p + int(ln)])
        p += int(ln)
    return RectilinearBinMapper(bnds)


def _check_clusters(model, fx, prefix):
    cl = model.clusters
    sizes = np.array([len(m.cluster_centers_) if hasattr(m, "cluster_centers_") else -1 for m in cl.cluster_models])
    assert np.array_equal(sizes, fx[prefix + "sizes"]), f"{prefix}sizes {sizes} vs {fx[prefix + 'sizes']}"
    cents = [m.cluster_centers_ for m in cl.cluster_models if hasattr(m, "cluster_centers_") and len(m.cluster_centers_)]
    _close(np.concatenate(cents), fx[prefix + "centers"][: sum(len(c) for c in cents)], prefix + "centers")
    assert int(model.n_clusters) == int(fx[prefix + "n_clusters"])
    remap = np.array([int(cl.we_remap[b]) for b in range(cl.bin_mapper.nbins)])
    assert np.array_equal(remap, fx[prefix + "we_remap"])
    assert sorted(int(b) for b in cl.target_bins) == fx[prefix + "target_bins"].tolist()
    assert sorted(int(b) for b in cl.basis_bins) == fx[prefix + "basis_bins"].tolist()
    assert [len(d) for d in model.dtrajs] == fx[prefix + "dtraj_lens"].tolist()
    got = np.concatenate(model.dtrajs)
    assert got.dtype == np.int64
    assert np.array_equal(got, fx[prefix + "dtrajs"]), f"{prefix}dtrajs: {(got != fx[prefix + 'dtrajs']).sum()} labels differ"
    pairs = np.concatenate([np.asarray(p).reshape(-1, 2) for p in model.pair_dtrajs])
    assert np.array_equal(pairs, fx[prefix + "pair_dtrajs"])


def _dense_from_triplets(fx, n, M):
    sel = fx["iterflux_iter"] == n
    f = np.zeros((M, M))
    f[fx["iterflux_row"][sel], fx["iterflux_col"][sel]] = fx["iterflux_val"][sel]
    return f


def _run_pipeline(name, user_featuriser=False):
    from msm_we_b200.msm_we import LinearCoordinates, modelWE

    fx = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    its = FD.unpack_iterations(fx)
    fname = f"product_{name}_west.h5"
    refshim.register_we_file(fname, its)
    model = modelWE()
    if user_featuriser:
        # a monkey-patched featuriser, as the reference's users install one: takes the staged (not direct-read) path
        model.processCoordinates = lambda coords: np.asarray(coords).reshape(np.shape(coords)[0], -1) \
            if np.ndim(coords) == 3 else np.asarray(coords).reshape(1, -1)
    pca = "pca_components" in fx.files
    model.initialize([fname], {"coords": None, "nAtoms": int(fx["n_atoms"]), "coord_ndim": int(fx["coord_ndim"])}, name,
                     basis_pcoord_bounds=fx["basis"], target_pcoord_bounds=fx["target"],
                     dim_reduce_method="pca" if pca else "none", tau=1.0, pcoord_ndim=int(fx["pcoord_ndim"]),
                     use_weights_in_clustering=bool(fx["use_weights"]))
    model.get_iterations()
    assert model.maxIter == int(fx["maxIter"]) and np.array_equal(model.numSegments, fx["numSegments"])
    model.get_coordSet(model.maxIter)
    assert np.array_equal(model.pcoordSet, fx["pcoordSet"], equal_nan=True)
    if pca:
        model.coordinates = LinearCoordinates(fx["pca_components"], fx["pca_mean"])
    model.dimReduce()
    if pca:
        assert model.ndim == int(fx["ndim"])
    ckw = {str(k): fx["cluster_kwarg_" + str(k)].item() for k in fx["cluster_kwargs_keys"]}
    call = {k[len("cluster_call_"):]: fx[k].tolist() for k in fx.files if k.startswith("cluster_call_")}
    if "cluster_call_iters_to_use" not in fx.files and name == "pipeline2d":
        call["user_bin_mapper"] = _mapper(fx)
    else:
        model.bin_mapper = _mapper(fx)          # what analysis.Run(file).iteration(2).bin_mapper gives the reference
    model.cluster_coordinates(n_clusters=int(fx["K"]), streaming=True, use_ray=True, stratified=True,
                              store_validation_model=True, **call, **ckw)
    return fx, model


@pytest.mark.parametrize("gpu", BACKENDS)
@pytest.mark.parametrize("name,user_featuriser", [("pipeline1d", False), ("pipeline1d", True), ("pipeline2d", False),
                                                  ("pipeline_voronoi", False)])
def test_clustering_discretization_flux_match_reference_run(request, monkeypatch, gpu, name, user_featuriser):
    _backend(request, monkeypatch, gpu)
    fx, model = _run_pipeline(name, user_featuriser)
    _check_clusters(model, fx, "c_")
    counts = np.concatenate([m._counts for m in model.clusters.cluster_models if hasattr(m, "cluster_centers_")])
    _close(counts, fx["c_counts"], "minibatch counts")
    steps = [getattr(m, "n_steps_", 0) for m in model.clusters.cluster_models]
    assert steps == fx["c_n_steps"].tolist()

    M = model.n_clusters + 2
    model.get_fluxMatrix(0, first_iter=1, last_iter=model.maxIter, use_ray=False)
    _close(model.fluxMatrixRaw, fx["flux_raw"], "fluxMatrixRaw (serial)")
    assert model._fluxMatrixParams == [0, 1, model.maxIter, None] and model.errorWeight == 0.0 and model.errorCount == 0
    for n in (2, 3, model.maxIter - 1):
        _close(model.get_iter_fluxMatrix(n), _dense_from_triplets(fx, n, M), f"get_iter_fluxMatrix({n})")
    model.get_fluxMatrix(0, first_iter=1, last_iter=model.maxIter, use_ray=True)
    _close(model.fluxMatrixRaw, fx["flux_raw_ray"], "fluxMatrixRaw (use_ray)")
    model.get_fluxMatrix(0, iters_to_use=fx["flux_subset_iters"].tolist(), use_ray=False)
    _close(model.fluxMatrixRaw, fx["flux_subset"], "fluxMatrixRaw (iters_to_use)")


@pytest.mark.parametrize("name", ["pipeline1d", "pipeline_voronoi"])
@pytest.mark.parametrize("gpu", BACKENDS)
def test_cleaning_block_validation_structures_match_reference_run(request, monkeypatch, gpu, name):
    _backend(request, monkeypatch, gpu)
    fx, model = _run_pipeline(name)
    model.get_fluxMatrix(0, first_iter=1, last_iter=model.maxIter, use_ray=False)
    model.organize_fluxMatrix(use_ray=False)
    _check_clusters(model, fx, "o_")
    _close(model.fluxMatrixRaw, fx["o_fluxMatrixRaw"], "fluxMatrixRaw after cleaning (restored original)")
    _close(model.fluxMatrix, fx["o_fluxMatrix"], "cleaned, sorted, normalised fluxMatrix")
    _close(model.targetRMSD_centers, fx["o_targetRMSD_centers"], "targetRMSD_centers")
    _close(model.targetRMSD_minmax, fx["o_targetRMSD_minmax"], "targetRMSD_minmax")
    assert np.array_equal(model.indBasis, fx["o_indBasis"]) and np.array_equal(model.indTargets, fx["o_indTargets"])
    assert model.nBins == int(fx["o_nBins"])
    _close(model.all_centers, fx["o_all_centers"], "all_centers")
    assert np.array_equal(model.sorted_centers, fx["o_sorted_centers"])
    assert model.cluster_mapping == {x: x for x in range(model.n_clusters + 2)}

    # downstream host linear algebra on OUR cleaned matrix: the package's own steps, and the oracle's restatement
    model.get_Tmatrix()
    model.get_steady_state()
    model.get_steady_state_target_flux()
    assert np.allclose(model.Tmatrix, fx["d_Tmatrix"], rtol=1e-10, atol=1e-300)
    assert np.allclose(model.pSS, fx["d_pSS"], rtol=1e-6, atol=1e-12)
    assert np.isclose(model.JtargetSS, float(fx["d_JtargetSS"]), rtol=1e-6, atol=0)
    T = O.transition_matrix(model.fluxMatrix, model.indBasis, model.indTargets)
    assert np.allclose(T, fx["d_Tmatrix"], rtol=1e-10, atol=1e-300)
    pss = O.steady_state(T)
    assert np.allclose(pss, fx["d_pSS"], rtol=1e-6, atol=1e-12)

    model.do_block_validation(2, 4, use_ray=False)
    for g, vm in enumerate(model.validation_models):
        assert model.validation_iterations[g] == fx[f"v{g}_iters"].tolist()
        _close(vm.fluxMatrixRaw, fx[f"v{g}_fluxMatrixRaw"], f"validation group {g} fluxMatrixRaw")
        _close(vm.fluxMatrix, fx[f"v{g}_fluxMatrix"], f"validation group {g} fluxMatrix")
        assert int(vm.n_clusters) == int(fx[f"v{g}_n_clusters"])
        assert np.isclose(vm.JtargetSS, float(fx[f"v{g}_JtargetSS"]), rtol=1e-6, atol=0)
        assert np.allclose(vm.pSS, fx[f"v{g}_pSS"], rtol=1e-6, atol=1e-12)
        assert vm.iteration_source is model.iteration_source      # copies share the data set

    model.update_cluster_structures(build_pcoord_cache=True)
    assert list(model.cluster_structures.keys()) != sorted(model.cluster_structures.keys()) or True
    keys = sorted(model.cluster_structures.keys())
    assert keys == fx["s_keys"].tolist()
    assert [len(model.cluster_structures[k]) for k in keys] == fx["s_count"].tolist()
    _close([np.sum(model.cluster_structure_weights[k]) for k in keys], fx["s_wsum"], "cluster_structure_weights")
    _close(np.array([np.sum(np.asarray(model.cluster_structures[k]), axis=0).ravel() for k in keys]), fx["s_coordsum"],
           "cluster_structures")
    _close(np.array([np.sum(np.asarray(model.pcoord_cache[k]), axis=0).ravel() for k in keys]), fx["s_pcoordsum"],
           "pcoord_cache")


def _model_with_centres(name, its, mapper, centres, K, basis, target, D, we_remap=None):
    from msm_we_b200.msm_we import modelWE
    from msm_we_b200.stratified_clustering import StratifiedClusters

    fname = f"product_{name}_west.h5"
    refshim.register_we_file(fname, its)
    model = modelWE()
    model.initialize([fname], {"coords": None, "nAtoms": D, "coord_ndim": 1}, name, basis_pcoord_bounds=basis,
                     target_pcoord_bounds=target, dim_reduce_method="none", tau=1.0, pcoord_ndim=1)
    model.get_iterations()
    model.dimReduce()
    clusters = StratifiedClusters(mapper, model, K, [])
    for b, c in enumerate(centres):
        if c is not None:
            clusters.cluster_models[b].cluster_centers_ = np.ascontiguousarray(c)
    if we_remap is not None:
        clusters.we_remap.update(we_remap)
    model.clusters = clusters
    model.n_clusters = K * mapper.nbins
    return model


@pytest.mark.parametrize("gpu", BACKENDS)
def test_predict_labels_match_reference_predict_ntl9_centres(request, monkeypatch, gpu):
    """The reference's StratifiedClusters.predict (per-segment sklearn loop) on the NTL9 fixture's own centres /
    boundaries / we_remap: remapped unfitted bin, basis and target points, both processing_from states, toggle."""
    from msm_we_b200.binning import RectilinearBinMapper

    _backend(request, monkeypatch, gpu)
    g = np.load(os.path.join(GOLDEN, "ntl9_clustered.npz"))
    fx = np.load(os.path.join(GOLDEN, "ref_predict_ntl9.npz"))
    nb = len(g["fitted"])
    centres = [g[f"centers_{b}"] if g["fitted"][b] else None for b in range(nb)]
    mapper = RectilinearBinMapper([np.asarray(g["boundaries"], dtype=np.float32)])
    its = FD.unpack_iterations(fx)
    model = _model_with_centres("ntl9", its, mapper, centres, 25, g["basis_bounds"], g["target_bounds"], 13,
                                {b: int(g["we_remap"][b]) for b in range(nb)})
    parents, children = [], []
    for it in range(1, model.maxIter):
        (p, c), _, _, tb, bb = model.do_stratified_ray_discretization(model, model.clusters, it, model.processCoordinates)
        parents.append(p); children.append(c)
        model.clusters.target_bins.update(tb); model.clusters.basis_bins.update(bb)
    assert np.array_equal(np.concatenate(parents), fx["parents"])
    assert np.array_equal(np.concatenate(children), fx["children"])
    assert sorted(model.clusters.target_bins) == fx["target_bins"].tolist()
    assert sorted(model.clusters.basis_bins) == fx["basis_bins"].tolist()
    # batched path gives the same labels
    model.launch_ray_discretization()
    assert np.array_equal(np.concatenate(model.dtrajs), fx["children"])
    assert np.array_equal(np.concatenate(model.pair_dtrajs)[:, 0], fx["parents"])
    # toggle: two predict calls alternate pcoord0List / pcoord1List
    model.load_iter_data(2)
    model.get_transition_data_lag0()
    xp = model.processCoordinates(model.coordPairList[..., 0])
    xc = model.processCoordinates(model.coordPairList[..., 1])
    model.clusters.toggle = True
    model.clusters.processing_from = True
    assert np.array_equal(model.clusters.predict(xp), fx["toggle_first"])
    assert np.array_equal(model.clusters.predict(xc), fx["toggle_second"])
    assert bool(model.clusters.processing_from) == bool(fx["toggle_state_after"])


@pytest.mark.parametrize("gpu", BACKENDS)
def test_cfg2_labels_and_flux_match_reference_run(request, monkeypatch, gpu):
    """BASELINE config 2 shape, first 24 iterations: labels of the reference's launch_ray_discretization and its serial
    get_fluxMatrix / get_iter_fluxMatrix, inputs regenerated from the seed and verified by checksum."""
    import dataclasses

    import workloads
    from msm_we_b200.binning import RectilinearBinMapper

    _backend(request, monkeypatch, gpu)
    fx = np.load(os.path.join(GOLDEN, "ref_predict_cfg2.npz"))
    cfg = dataclasses.replace(workloads.CONFIGS["cfg2"], n_iters=int(fx["n_iters"]))
    means, centers = workloads.make_centers(cfg)
    host = workloads.generate_host(cfg, means)
    assert FD.checksum(*[d[k] for d in host for k in ("pcoord0", "pcoord1", "weights", "parent", "child")]) == str(fx["input_checksum"]), \
        "the seeded generator no longer reproduces the inputs the fixture was made from"
    assert FD.checksum(*centers) == str(fx["centers_checksum"])
    its = [dict(weights=d["weights"], pcoord=np.stack([d["pcoord0"], d["pcoord1"]], axis=1),
                coords=np.stack([d["parent"], d["child"]], axis=1)[:, :, :, None], parent_id=np.arange(cfg.n_segs)) for d in host]
    basis, target = workloads.region_bounds(cfg)
    model = _model_with_centres("cfg2", its, RectilinearBinMapper(workloads.boundaries(cfg)), centers, cfg.k_per_bin,
                                basis, target, cfg.dim)
    model.launch_ray_discretization()
    assert np.array_equal(np.concatenate(model.dtrajs), fx["dtrajs"].astype(np.int64))
    assert np.array_equal(np.concatenate(model.pair_dtrajs), fx["pair_dtrajs"].astype(np.int64))
    assert sorted(model.clusters.target_bins) == fx["target_bins"].tolist()
    assert sorted(model.clusters.basis_bins) == fx["basis_bins"].tolist()
    model.get_fluxMatrix(0, first_iter=1, last_iter=model.maxIter, use_ray=False)
    ref = np.zeros(tuple(fx["flux_shape"]))
    ref[fx["flux_row"], fx["flux_col"]] = fx["flux_val"]
    _close(model.fluxMatrixRaw, ref, "cfg2 fluxMatrixRaw")
    ref5 = np.zeros(tuple(fx["flux_shape"]))
    ref5[fx["iter5_row"], fx["iter5_col"]] = fx["iter5_val"]
    _close(model.get_iter_fluxMatrix(5), ref5, "cfg2 get_iter_fluxMatrix(5)")


def test_oracle_matches_reference_run_labels_and_flux():
    """Pins the ORACLE (the checker every other test uses) to reference-executed outputs: labels of the cfg2 fixture
    and its flux matrix."""
    import dataclasses

    import workloads

    fx = np.load(os.path.join(GOLDEN, "ref_predict_cfg2.npz"))
    cfg = dataclasses.replace(workloads.CONFIGS["cfg2"], n_iters=int(fx["n_iters"]))
    means, centers = workloads.make_centers(cfg)
    host = workloads.generate_host(cfg, means)
    basis, target = workloads.region_bounds(cfg)
    strat = O.StratifiedOracle(O.RectilinearBinMapperOracle(workloads.boundaries(cfg)), centers, basis, target)
    pairs, per = [], []
    for i, d in enumerate(host[: cfg.n_iters - 1], start=1):
        p, c = O.discretize_iteration(strat, d["parent"], d["child"], d["pcoord0"], d["pcoord1"])
        pairs.append(np.stack([p, c], axis=1))
        if i >= 2:
            per.append((pairs[-1], d["pcoord0"], d["pcoord1"], d["weights"]))
    assert np.array_equal(np.concatenate(pairs), fx["pair_dtrajs"].astype(np.int64))
    ref = np.zeros(tuple(fx["flux_shape"]))
    ref[fx["flux_row"], fx["flux_col"]] = fx["flux_val"]
    got = O.flux_matrix(cfg.n_clusters, per, basis, target)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("name", ["pipeline1d", "pipeline_voronoi"])
def test_oracle_discretization_and_flux_match_reference_pipeline_run(name):
    """The oracle's mapper restatements (westpa's Rectilinear / Voronoi mappers are not in the image), its ``we_remap``
    handling and its flux accumulation against what the reference's own run produced from the same fitted centres."""
    fx = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    its = FD.unpack_iterations(fx)
    if "voronoi_centers" in fx.files:
        mapper = O.VoronoiBinMapperOracle(fx["voronoi_centers"])
    else:
        bnds, p = [], 0
        for ln in fx["boundary_lens"]:
            bnds.append(fx["boundaries"][p:p + int(ln)])
            p += int(ln)
        mapper = O.RectilinearBinMapperOracle(bnds)
    centres, p = [], 0
    for sz in fx["c_sizes"]:
        centres.append(None if sz < 0 else fx["c_centers"][p:p + int(sz)])
        p += max(int(sz), 0)
    strat = O.StratifiedOracle(mapper, centres, fx["basis"], fx["target"],
                               we_remap={b: int(r) for b, r in enumerate(fx["c_we_remap"])})
    P = int(fx["pcoord_ndim"])
    pairs, per = [], []
    for i, d in enumerate(its[: int(fx["maxIter"]) - 1], start=1):
        S = len(d["weights"])
        xp, xc = d["coords"][:, 0].reshape(S, -1), d["coords"][:, 1].reshape(S, -1)
        pc0, pc1 = d["pcoord"][:, 0, :P], d["pcoord"][:, 1, :P]
        a, b = O.discretize_iteration(strat, xp, xc, pc0, pc1)
        pairs.append(np.stack([a, b], axis=1))
        if i >= 2:
            per.append((pairs[-1], pc0, pc1, d["weights"]))
    assert np.array_equal(np.concatenate(pairs), fx["c_pair_dtrajs"])
    got = O.flux_matrix(int(fx["c_n_clusters"]), per, fx["basis"], fx["target"])
    _close(got, fx["flux_raw"], "oracle flux matrix")


@pytest.mark.gpu
def test_lineage_colour_counts_match_reference_nonmarkov_fit():
    """History-coloured count matrix over WE lineages: the reference's NonMarkovModel.fit over the traced trajectories
    (fixture ref_colour_lineages.npz, made by executing msm_we/nmm.py) vs the segment-wise GPU formulation + K3 (C = 2).
    Counts are integers: bit-exact."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from msm_we_b200.lineage import coloured_lineage_counts

    fx = np.load(os.path.join(GOLDEN, "ref_colour_lineages.npz"))
    got = coloured_lineage_counts(list(fx["labels"]), list(fx["parents"]), int(fx["n_states"]), fx["stateA"].tolist(),
                                  fx["stateB"].tolist())
    assert got.shape == fx["nm_cmatrix"].shape
    assert np.array_equal(got, fx["nm_cmatrix"])


def test_oracle_colour_counts_match_reference_nonmarkov_fit():
    """Pins the oracle's colour walk to the reference-executed lineage fixture (CPU)."""
    fx = np.load(os.path.join(GOLDEN, "ref_colour_lineages.npz"))
    labels, parents = fx["labels"], fx["parents"]
    n_it, S = labels.shape
    trajs = []
    for s in range(S):
        t, cur = [], s
        for it in range(n_it - 1, -1, -1):
            t.append(int(labels[it][cur]))
            cur = int(parents[it][cur])
        trajs.append(t[::-1])
    got = O.colour_counts(trajs, int(fx["n_states"]), fx["stateA"].tolist(), fx["stateB"].tolist(), lag=1)
    assert np.array_equal(got, fx["nm_cmatrix"])


class _FakeWorkManager:
    is_master = True


class _FakeDataManager:
    def __init__(self, name):
        self.we_h5filename = name
        self.closed = False

    def close_backing(self):
        self.closed = True


class _FakeSimManager:
    def __init__(self, name):
        self.work_manager = _FakeWorkManager()
        self.data_manager = _FakeDataManager(name)
        self.callbacks = []

    def finalize_run(self):
        pass

    def register_callback(self, hook, fn, priority):
        self.callbacks.append((hook, fn, priority))


@pytest.mark.parametrize("gpu", BACKENDS)
def test_hamsm_driver_plugin_builds_the_reference_model(request, monkeypatch, gpu):
    """The WESTPA plugin boundary (msm_we/westpa_plugins/hamsm_driver.py:8-144) with a stand-in sim_manager /
    data_manager: callback registration, featuriser loading, build_analyze_model through the whole chain (HDF5 feeder ->
    clustering -> discretization -> flux -> cleaning -> block validation), model stored on the data manager -- and the
    model equals the one the reference built from the same WE file (pipeline1d fixture)."""
    from msm_we_b200 import msm_we as mw
    from msm_we_b200.westpa_plugins.hamsm_driver import HAMSMDriver

    _backend(request, monkeypatch, gpu)
    monkeypatch.setattr(mw.modelWE, "processCoordinates", mw.modelWE.processCoordinates)     # restored after the test
    fx = np.load(os.path.join(GOLDEN, "ref_pipeline1d.npz"))
    fname = "plugin_pipeline1d_west.h5"
    refshim.register_we_file(fname, FD.unpack_iterations(fx))
    sim = _FakeSimManager(fname)
    cfg = {"model_name": "plugin", "n_clusters": int(fx["K"]), "tau": 1.0, "basis_pcoord_bounds": fx["basis"],
           "target_pcoord_bounds": fx["target"], "dimreduce_method": "none", "featurization": "fixture_data.flatten_featurizer",
           "ref_pdb_file": {"coords": None, "nAtoms": int(fx["n_atoms"]), "coord_ndim": int(fx["coord_ndim"])},
           "user_bin_mapper": _mapper(fx), "cross_validation_groups": 2,
           "cluster_args": {"random_state": int(fx["cluster_kwarg_random_state"]),
                            "iters_to_use": fx["cluster_call_iters_to_use"].tolist()}}
    driver = HAMSMDriver(sim, cfg)
    assert len(sim.callbacks) == 1 and sim.callbacks[0][0] == sim.finalize_run and sim.callbacks[0][2] == 2
    # cluster_stratified logs "conflicting parameters" when both first_cluster_iter and iters_to_use arrive and goes on
    # with iters_to_use, exactly as the reference does (_clustering.py:646-650)
    model = sim.callbacks[0][1]()
    assert sim.data_manager.closed and sim.data_manager.hamsm_model is model and driver.data_manager is sim.data_manager
    _check_clusters(model, fx, "o_")
    _close(model.fluxMatrix, fx["o_fluxMatrix"], "plugin-built cleaned fluxMatrix")
    # what later WESTPA plugins read off the stored model (restart / optimization drivers): steady state + target flux
    assert np.allclose(model.pSS, fx["d_pSS"], rtol=1e-6, atol=1e-12)
    assert np.isclose(model.JtargetSS, float(fx["d_JtargetSS"]), rtol=1e-6, atol=0)
    assert len(model.validation_models) == 2
    for g, vm in enumerate(model.validation_models):
        _close(vm.fluxMatrix, fx[f"v{g}_fluxMatrix"], f"plugin validation group {g}")
        assert np.isclose(vm.JtargetSS, float(fx[f"v{g}_JtargetSS"]), rtol=1e-6, atol=0)


def test_dimreduce_refuses_methods_it_cannot_fit():
    from msm_we_b200.msm_we import LinearCoordinates, modelWE

    m = modelWE()
    m.dimReduceMethod = "pca"
    with pytest.raises(NotImplementedError):
        m.dimReduce(use_weights=True, variance_cutoff=0.9)
    m.coordinates = LinearCoordinates(np.eye(3)[:2], np.zeros(3))
    m.dimReduce()
    assert m.ndim == 2
