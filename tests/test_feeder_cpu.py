"""WESTPA HDF5 feeder (reference msm_we/_hamsm/_data.py:254-320, 531-555, 807-993) against what the
REFERENCE's own loader left on its model for a run spread over two west.h5 files (fixture ref_feeder_twofiles.npz, made by
tests/golden/make_reference_fixtures.py feeder), plus the source's own contracts.  h5py is not in the image: both sides
read the same in-memory stand-in of the WESTPA layout (tests/golden/refshim.py)."""
import os
import sys

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)

import fixture_data as FD  # noqa: E402
import refshim  # noqa: E402


@pytest.fixture()
def two_files(monkeypatch):
    from cpu_emulation import emulate_kernels

    emulate_kernels(monkeypatch)
    monkeypatch.setitem(sys.modules, "h5py", refshim.fake_h5py_module())
    fx = np.load(os.path.join(GOLDEN, "ref_feeder_twofiles.npz"))
    its = FD.unpack_iterations(fx)
    part = lambda d, sl: {k: v[sl] for k, v in d.items()}          # noqa: E731
    refshim.register_we_file("p_feeder_a_west.h5", [part(d, slice(0, int(c))) for d, c in zip(its, fx["cut"])])
    refshim.register_we_file("p_feeder_b_west.h5", [part(d, slice(int(c), None)) for d, c in zip(its, fx["cut"])])
    return fx, its, ["p_feeder_a_west.h5", "p_feeder_b_west.h5"]


def _model(files):
    from msm_we_b200.msm_we import modelWE

    model = modelWE()
    model.initialize(files, {"coords": None, "nAtoms": 3, "coord_ndim": 3}, "feeder", basis_pcoord_bounds=[[0.0, 0.5]],
                     target_pcoord_bounds=[[3.5, 1.0e6]], dim_reduce_method="none", tau=1.0, pcoord_ndim=1)
    model.get_iterations()
    return model


def test_loader_matches_reference_loader_on_a_two_file_run(two_files):
    fx, its, files = two_files
    model = _model(files)
    assert model.maxIter == int(fx["maxIter"]) and np.array_equal(model.numSegments, fx["numSegments"])
    offs = np.concatenate([[0], np.cumsum(fx["f_lens"])])
    for n in range(1, model.maxIter):
        a, b = offs[n - 1], offs[n]
        model.load_iter_data(n)
        assert model.nSeg == b - a
        assert np.array_equal(model.weightList, fx["f_weights"][a:b])
        assert np.array_equal(model.pcoord0List, fx["f_pcoord0"][a:b]) and model.pcoord0List.shape == (b - a, 1)
        assert np.array_equal(model.pcoord1List, fx["f_pcoord1"][a:b])
        assert np.array_equal(model.westList, fx["f_west"][a:b]) and np.array_equal(model.segindList, fx["f_segind"][a:b])
        assert np.array_equal(model.get_iter_coordinates(n), fx["f_child"][a:b])
        model.get_transition_data_lag0()
        assert np.array_equal(model.coordPairList, fx["f_pairs"][a:b])
        assert np.array_equal(model.transitionWeights, fx["f_weights"][a:b])


def test_source_contracts(two_files):
    from msm_we_b200._hamsm._data import H5IterationSource

    fx, its, files = two_files
    src = H5IterationSource(files, pcoord_ndim=1)
    assert src.n_iterations() == int(fx["maxIter"])            # (the appended last group holds no dynamics)
    assert not src.has(int(fx["maxIter"]) + 1) and src.has(1)
    for n in (1, 3):
        rec = src.get(n)
        S = rec.weights.shape[0]
        assert S == src.n_segments(n) == len(its[n - 1]["weights"])
        # segments of the first file, then of the second; indices restart in every file
        c = int(fx["cut"][n - 1])
        assert np.array_equal(rec.west_file, np.r_[np.zeros(c, int), np.ones(S - c, int)])
        assert np.array_equal(rec.seg_index, np.r_[np.arange(c), np.arange(S - c)])
        assert np.array_equal(rec.parent_id, its[n - 1]["parent_id"])
        # straight into caller-owned rows == through get()
        dp, dc = np.empty((S, 9)), np.empty((S, 9))
        assert src.read_pair_into(n, dp, dc)
        assert np.array_equal(dp, rec.parent_coords.reshape(S, -1)) and np.array_equal(dc, rec.child_coords.reshape(S, -1))
        light = src.get(n, coords=False)
        assert np.array_equal(light.weights, rec.weights)
    import copy, pickle
    assert copy.deepcopy(src) is src                           # copies of the model share the source
    again = pickle.loads(pickle.dumps(src))
    assert again.file_list == src.file_list and again.n_segments(2) == src.n_segments(2)


def test_source_errors(monkeypatch):
    from msm_we_b200._hamsm._data import H5IterationSource

    monkeypatch.setitem(sys.modules, "h5py", refshim.fake_h5py_module())
    its = FD.we_dataset(seed=5, n_iters=3, segs0=6, seg_growth=0, n_atoms=2, coord_ndim=3, bins_per_dim=4, k_true=1)
    refshim.register_we_file("p_err_west.h5", its)
    src = H5IterationSource(["p_err_west.h5"], auxpath="not_there")
    src.get(1, coords=False)                                   # pcoords and weights do not need the structures
    with pytest.raises(KeyError):
        src.get(1)                                             # the run was not augmented with this dataset
    single = [dict(d, coords=d["coords"][:, :1]) for d in its]
    refshim.register_we_file("p_single_west.h5", single)
    with pytest.raises(AssertionError):
        H5IterationSource(["p_single_west.h5"]).get(1)          # start AND end structure are needed for a transition
    with pytest.raises(KeyError):
        H5IterationSource(["p_err_west.h5"]).get(17)


BACKENDS = [pytest.param(False, id="host-logic-cpu"), pytest.param(True, id="cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("gpu", BACKENDS)
def test_nan_frames_direct_read_path_equals_staged_path(monkeypatch, gpu):
    """A broken (NaN) structure: the reference zeroes that segment's weight in the flux pass (_data.py:302-313).  The
    direct-read discretization finds such rows on the device (rows_with_nan) and hands them to the flux pass; the
    staged path finds them on the host.  Both must give the same labels for the intact frames and the same flux matrix,
    and the broken segments must not contribute."""
    import torch

    if gpu and not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    if not gpu:
        from cpu_emulation import emulate_kernels

        emulate_kernels(monkeypatch)
    monkeypatch.setitem(sys.modules, "h5py", refshim.fake_h5py_module())
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.msm_we import modelWE
    from msm_we_b200.stratified_clustering import StratifiedClusters

    its = FD.we_dataset(seed=77, n_iters=7, segs0=60, seg_growth=2, n_atoms=3, coord_ndim=3, bins_per_dim=4, k_true=2)
    broken = {2: [5, 17], 4: [0], 5: [33]}                       # iteration -> segments with a NaN end or start structure
    for it, segs in broken.items():
        for s in segs:
            its[it - 1]["coords"][s, (s % 2), 1, 2] = np.nan      # start structure for odd s, end structure for even s
    refshim.register_we_file("p_nan_west.h5", its)
    rng = np.random.default_rng(3)
    centres = [rng.normal(0, 3, size=(3, 9)) for _ in range(4)]
    results = []
    for user_featuriser in (False, True):
        model = modelWE()
        if user_featuriser:                                       # a monkey-patched featuriser: the staged path
            model.processCoordinates = lambda c: np.asarray(c).reshape(np.shape(c)[0], -1)
        model.initialize(["p_nan_west.h5"], {"coords": None, "nAtoms": 3, "coord_ndim": 3}, "nan",
                         basis_pcoord_bounds=[[0.0, 0.5]], target_pcoord_bounds=[[3.5, 1.0e6]], dim_reduce_method="none",
                         tau=1.0, pcoord_ndim=1)
        model.get_iterations()
        model.dimReduce()
        clusters = StratifiedClusters(RectilinearBinMapper(FD.boundaries(4)), model, 3, [])
        for b in range(4):
            clusters.cluster_models[b].cluster_centers_ = centres[b].copy()
        model.clusters = clusters
        model.n_clusters = 12
        model.launch_ray_discretization()
        model.get_fluxMatrix(0, first_iter=1, last_iter=model.maxIter)
        results.append((np.concatenate(model.pair_dtrajs), model.fluxMatrixRaw.copy(), model))
    (pa, fa, ma), (pb, fb, mb) = results
    offs = np.concatenate([[0], np.cumsum([len(d["weights"]) for d in its])])
    bad = np.zeros(len(pa), dtype=bool)
    for it, segs in broken.items():
        bad[offs[it - 1] + np.array(segs)] = True
    assert np.array_equal(pa[~bad[: len(pa)]], pb[~bad[: len(pb)]])
    assert np.array_equal(fa, fb)
    # the broken segments contribute nothing: total weight in the matrix == weight of the intact segments / nI
    used = range(2, ma.maxIter)                                   # get_fluxMatrix covers first_iter + 1 .. last_iter - 1
    total = sum(its[n - 1]["weights"][[s for s in range(len(its[n - 1]["weights"])) if s not in broken.get(n, [])]].sum()
                for n in used)
    assert abs(fa.sum() - total / len(used)) <= 1e-12 * total


@pytest.mark.parametrize("gpu", BACKENDS)
def test_lloyd_refinement_from_hdf5_source_equals_array_source(monkeypatch, gpu):
    """lloyd_refine_clusters stages rows of a source that does not own arrays (HDF5) through pinned buffers; same
    refined centres as with the in-memory source, and the discretization that follows (direct-read path) is unaffected
    by the rows the refinement left on the device."""
    import torch

    if gpu and not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    if not gpu:
        from cpu_emulation import emulate_kernels

        emulate_kernels(monkeypatch)
    monkeypatch.setitem(sys.modules, "h5py", refshim.fake_h5py_module())
    from msm_we_b200._hamsm._data import ArrayIterationSource, IterationRecord
    from msm_we_b200.binning import RectilinearBinMapper
    from msm_we_b200.msm_we import modelWE
    from msm_we_b200.stratified_clustering import StratifiedClusters

    its = FD.we_dataset(seed=91, n_iters=8, segs0=90, seg_growth=0, n_atoms=3, coord_ndim=3, bins_per_dim=4, k_true=2)
    refshim.register_we_file("p_lloyd_west.h5", its)
    arr = ArrayIterationSource()
    for i, d in enumerate(its, start=1):
        S = len(d["weights"])
        arr.add(i, IterationRecord(d["pcoord"][:, 0], d["pcoord"][:, 1], d["weights"], d["coords"][:, 0].reshape(S, -1),
                                   d["coords"][:, 1].reshape(S, -1)))
    rng = np.random.default_rng(4)
    centres = [rng.normal(0, 3, size=(3, 9)) for _ in range(4)]
    out = []
    for source in (["p_lloyd_west.h5"], arr):
        model = modelWE()
        model.initialize(source, {"coords": None, "nAtoms": 3, "coord_ndim": 3}, "lloyd", basis_pcoord_bounds=[[0.0, 0.5]],
                         target_pcoord_bounds=[[3.5, 1.0e6]], dim_reduce_method="none", tau=1.0, pcoord_ndim=1)
        model.get_iterations()
        model.dimReduce()
        clusters = StratifiedClusters(RectilinearBinMapper(FD.boundaries(4)), model, 3, [])
        for b in range(4):
            clusters.cluster_models[b].cluster_centers_ = centres[b].copy()
        model.clusters = clusters
        model.n_clusters = 12
        n = model.lloyd_refine_clusters(3)
        refined = np.concatenate([m.cluster_centers_ for m in clusters.cluster_models])
        model.launch_ray_discretization()
        assert model._resident_child_rows is None
        out.append((n, refined, np.concatenate(model.pair_dtrajs)))
    assert out[0][0] == out[1][0] > 0
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2])
