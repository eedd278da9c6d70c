"""Golden fixtures produced by EXECUTING the reference's own code (/root/reference/msm_we, unmodified) in this
container -- run here only; /root/reference does not exist on the GPU box, the .npz outputs are committed.

    OMP_NUM_THREADS=1 python tests/golden/make_reference_fixtures.py [pipeline1d pipeline2d pipeline_voronoi predict_cfg2 predict_ntl9 colour]

How the reference is made importable is described in tests/golden/refshim.py (stand-ins for ray / h5py / westpa /
mdtraj / deeptime / matplotlib, none of which holds arithmetic of the path; sklearn 1.9 / scipy 1.18 / numpy 2.3 are
the installed ones).  What runs is the reference's real call chain:

  pipeline*    modelWE.initialize -> get_iterations -> get_coordSet -> dimReduce -> cluster_coordinates
               (cluster_stratified -> do_stratified_clustering -> MiniBatchKMeans.partial_fit;
                launch_ray_discretization -> do_stratified_ray_discretization -> StratifiedClusters.predict)
               -> get_fluxMatrix / get_iter_fluxMatrix -> organize_fluxMatrix (organize_stratified, get_cluster_centers)
               -> get_Tmatrix -> get_steady_state -> get_steady_state_target_flux -> do_block_validation
               -> update_cluster_structures
               reference: msm_we/msm_we.py:588-1009, _hamsm/_clustering.py:142-195, 525-1142, 1144-1329, 1398-1611,
               _hamsm/_fluxmatrix.py:21-345, _hamsm/_data.py:254-320, 531-618, 677-759, 807-993
  predict_*    StratifiedClusters.predict with fixed centres, both ``processing_from`` states, remapped / unfitted WE
               bins, basis / target points; FluxMatrixMixin.get_iter_fluxMatrix and the serial get_fluxMatrix loop
               reference: msm_we/stratified_clustering.py:101-212, _hamsm/_fluxmatrix.py:21-72, 97-164, 232-260, 342
  colour       NonMarkovModel.fit count matrix over WE lineage trajectories (msm_we/nmm.py:117-167)
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import fixture_data as FD  # noqa: E402
import refshim  # noqa: E402


def processCoordinates(self, coords):
    """The featuriser users monkey-patch onto modelWE (docs/usage.rst:41-60): flatten [S, nAtoms, 3] -> [S, 3 nAtoms]."""
    coords = np.asarray(coords)
    if coords.ndim == 2:
        return coords.reshape(1, -1)
    return coords.reshape(coords.shape[0], -1)


def _quiet():
    import logging

    logging.getLogger("msm_we").setLevel(logging.ERROR)
    logging.getLogger("msm_we._logging").setLevel(logging.ERROR)


def _cat(list_of_arrays, dtype=None):
    if len(list_of_arrays) == 0:
        return np.zeros(0, dtype=dtype or np.float64)
    return np.concatenate([np.asarray(a, dtype=dtype) for a in list_of_arrays])


def snapshot_clusters(model, prefix):
    """Everything the clustering + discretization steps leave on the model."""
    out = {}
    cl = model.clusters
    sizes, cents, counts, steps = [], [], [], []
    for m in cl.cluster_models:
        if hasattr(m, "cluster_centers_"):
            sizes.append(len(m.cluster_centers_))
            cents.append(np.asarray(m.cluster_centers_, dtype=np.float64))
            counts.append(np.asarray(getattr(m, "_counts", np.zeros(len(m.cluster_centers_))), dtype=np.float64)
                          [: len(m.cluster_centers_)] if len(getattr(m, "_counts", [])) >= len(m.cluster_centers_)
                          else np.zeros(len(m.cluster_centers_)))
            steps.append(int(getattr(m, "n_steps_", 0)))
        else:
            sizes.append(-1)          # never fitted: no cluster_centers_ attribute
            steps.append(0)
    out[prefix + "sizes"] = np.array(sizes, dtype=np.int64)
    out[prefix + "centers"] = np.concatenate(cents, axis=0) if cents else np.zeros((0, 0))
    out[prefix + "counts"] = _cat(counts)
    out[prefix + "n_steps"] = np.array(steps, dtype=np.int64)
    out[prefix + "we_remap"] = np.array([int(cl.we_remap[b]) for b in range(cl.bin_mapper.nbins)], dtype=np.int64)
    out[prefix + "n_clusters"] = np.int64(model.n_clusters)
    out[prefix + "target_bins"] = np.array(sorted(int(b) for b in cl.target_bins), dtype=np.int64)
    out[prefix + "basis_bins"] = np.array(sorted(int(b) for b in cl.basis_bins), dtype=np.int64)
    out[prefix + "dtraj_lens"] = np.array([len(d) for d in model.dtrajs], dtype=np.int64)
    out[prefix + "dtrajs"] = _cat(model.dtrajs, np.int64)
    out[prefix + "pair_dtrajs"] = np.concatenate([np.asarray(p, dtype=np.int64).reshape(-1, 2) for p in model.pair_dtrajs])
    return out


def _euclid(coord, centers):
    """The distance function a WESTPA user hands to VoronoiBinMapper."""
    return np.linalg.norm(np.asarray(centers, dtype=np.float64) - np.asarray(coord, dtype=np.float64), axis=1)


def run_pipeline(name, its, bnds, basis, target, K, pcoord_ndim, n_atoms, coord_ndim, dim_reduce="none",
                 use_weights=False, cluster_kwargs=None, user_mapper=False, cluster_call_kwargs=None, organize=True,
                 voronoi_centers=None):
    msm_we = refshim.load_reference()
    from msm_we._hamsm import _clustering as RC

    refshim.ragged_numpy(RC)
    _quiet()
    fname = f"{name}_west.h5"
    if voronoi_centers is not None:
        vc = np.asarray(voronoi_centers, dtype=np.float64)
        mapper = refshim.VoronoiBinMapper(_euclid, vc)
        # no progress coordinate may sit on a cell boundary to within float32 resolution (WESTPA compares in float32)
        allpc = np.concatenate([d["pcoord"].reshape(-1, pcoord_ndim) for d in its])
        d = np.sort(np.stack([_euclid(c, vc) for c in allpc]), axis=1)
        assert (d[:, 1] - d[:, 0]).min() > 2e-5, (d[:, 1] - d[:, 0]).min()
    else:
        mapper = refshim.RectilinearBinMapper(bnds)
    refshim.register_we_file(fname, its, bin_mapper=mapper)
    msm_we.modelWE.processCoordinates = processCoordinates
    model = msm_we.modelWE()
    model.initialize([fname], {"coords": None, "nAtoms": n_atoms, "coord_ndim": coord_ndim}, name,
                     basis_pcoord_bounds=basis, target_pcoord_bounds=target, dim_reduce_method=dim_reduce, tau=1.0,
                     pcoord_ndim=pcoord_ndim, use_weights_in_clustering=use_weights)
    model.get_iterations()
    model.get_coordSet(model.maxIter)
    model.dimReduce()
    out = FD.pack_iterations(its)
    out.update(boundaries=np.concatenate(bnds), boundary_lens=np.array([len(b) for b in bnds], dtype=np.int64),
               basis=np.asarray(basis, dtype=np.float64), target=np.asarray(target, dtype=np.float64), K=np.int64(K),
               n_atoms=np.int64(n_atoms), coord_ndim=np.int64(coord_ndim), pcoord_ndim=np.int64(pcoord_ndim),
               use_weights=np.bool_(use_weights), maxIter=np.int64(model.maxIter),
               numSegments=np.asarray(model.numSegments, dtype=np.float64), pcoordSet=np.asarray(model.pcoordSet))
    if voronoi_centers is not None:
        out["voronoi_centers"] = np.asarray(voronoi_centers, dtype=np.float64)
    if dim_reduce == "pca":
        out["pca_components"] = np.asarray(model.coordinates.components_, dtype=np.float64)
        out["pca_mean"] = np.asarray(model.coordinates.mean_, dtype=np.float64)
        out["ndim"] = np.int64(model.ndim)

    ckw = dict(cluster_kwargs or {})
    call = dict(cluster_call_kwargs or {})
    if user_mapper:
        call["user_bin_mapper"] = mapper
    model.cluster_coordinates(n_clusters=K, streaming=True, use_ray=True, stratified=True, store_validation_model=True,
                              **call, **ckw)
    out["cluster_kwargs_keys"] = np.array(sorted(ckw.keys()))
    for k, v in ckw.items():
        out["cluster_kwarg_" + k] = np.asarray(v)
    for k, v in call.items():
        if k != "user_bin_mapper":
            out["cluster_call_" + k] = np.asarray(v)
    out.update(snapshot_clusters(model, "c_"))

    # flux matrix: serial path, then every per-iteration matrix (sparse triplets)
    model.get_fluxMatrix(0, first_iter=1, last_iter=model.maxIter, use_ray=False)
    out["flux_raw"] = model.fluxMatrixRaw.copy()
    rows, cols, vals, its_idx = [], [], [], []
    for n in range(2, model.maxIter):
        f = model.get_iter_fluxMatrix(n)
        r, c = np.nonzero(f)
        rows.append(r); cols.append(c); vals.append(f[r, c]); its_idx.append(np.full(len(r), n))
    out.update(iterflux_iter=_cat(its_idx, np.int64), iterflux_row=_cat(rows, np.int64), iterflux_col=_cat(cols, np.int64),
               iterflux_val=_cat(vals))
    # the Ray variant (synchronous stand-in => iteration order) and an explicit iteration subset
    model.get_fluxMatrix(0, first_iter=1, last_iter=model.maxIter, use_ray=True)
    out["flux_raw_ray"] = model.fluxMatrixRaw.copy()
    subset = list(range(3, model.maxIter, 2))
    model.get_fluxMatrix(0, iters_to_use=subset, use_ray=False)
    out["flux_subset_iters"] = np.array(subset, dtype=np.int64)
    out["flux_subset"] = model.fluxMatrixRaw.copy()
    model.get_fluxMatrix(0, first_iter=1, last_iter=model.maxIter, use_ray=False)

    if not organize:
        # the reference's cleaning step cannot run with a multi-dimensional pcoord
        # (update_sorted_cluster_centers, _clustering.py:1606, assigns a [P] vector to one element)
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB; clusters {int(out['c_n_clusters'])}, remap {out['c_we_remap'].tolist()}")
        return

    # cleaning: organize_fluxMatrix -> organize_stratified (re-discretize, get_cluster_centers, re-flux)
    model.organize_fluxMatrix(use_ray=False)
    out.update(snapshot_clusters(model, "o_"))
    out.update(o_fluxMatrix=model.fluxMatrix.copy(), o_fluxMatrixRaw=model.fluxMatrixRaw.copy(),
               o_targetRMSD_centers=np.asarray(model.targetRMSD_centers, dtype=np.float64),
               o_targetRMSD_minmax=np.asarray(model.targetRMSD_minmax, dtype=np.float64),
               o_indBasis=np.asarray(model.indBasis), o_indTargets=np.asarray(model.indTargets),
               o_nBins=np.int64(model.nBins), o_all_centers=np.asarray(model.all_centers, dtype=np.float64),
               o_sorted_centers=np.asarray(model.sorted_centers, dtype=np.int64))
    model.get_Tmatrix()
    model.get_steady_state()
    model.get_steady_state_target_flux()
    out.update(d_Tmatrix=np.asarray(model.Tmatrix), d_pSS=np.asarray(model.pSS).ravel(), d_JtargetSS=np.float64(model.JtargetSS))

    # block validation (msm_we.py:884-1009)
    model.do_block_validation(2, 4, use_ray=False)
    for g, vm in enumerate(model.validation_models):
        out[f"v{g}_iters"] = np.array(model.validation_iterations[g], dtype=np.int64)
        out[f"v{g}_fluxMatrixRaw"] = vm.fluxMatrixRaw.copy()
        out[f"v{g}_fluxMatrix"] = vm.fluxMatrix.copy()
        out[f"v{g}_n_clusters"] = np.int64(vm.n_clusters)
        out[f"v{g}_JtargetSS"] = np.float64(vm.JtargetSS)
        out[f"v{g}_pSS"] = np.asarray(vm.pSS).ravel()

    # cluster structures (_clustering.py:1398-1526): per cluster the member count, weight sum and coordinate sum
    model.update_cluster_structures(build_pcoord_cache=True)
    keys = sorted(int(k) for k in model.cluster_structures.keys())
    out["s_keys"] = np.array(keys, dtype=np.int64)
    out["s_count"] = np.array([len(model.cluster_structures[k]) for k in keys], dtype=np.int64)
    out["s_wsum"] = np.array([float(np.sum(model.cluster_structure_weights[k])) for k in keys])
    out["s_coordsum"] = np.array([np.sum(np.asarray(model.cluster_structures[k], dtype=np.float64), axis=0).ravel()
                                  for k in keys])
    out["s_pcoordsum"] = np.array([np.sum(np.asarray(model.pcoord_cache[k], dtype=np.float64), axis=0).ravel()
                                   for k in keys])
    path = os.path.join(HERE, f"ref_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB; clusters {int(out['c_n_clusters'])} -> {int(out['o_n_clusters'])}, "
          f"J={float(out['d_JtargetSS']):.3e}, remap {out['c_we_remap'].tolist()} -> {out['o_we_remap'].tolist()}")


# ---------------------------------------------------------------------------------------------------------------
def _model_with_centres(msm_we, name, its, mapper, centres_per_bin, K, basis, target, n_atoms, coord_ndim, pcoord_ndim,
                        we_remap=None):
    from msm_we.stratified_clustering import StratifiedClusters

    fname = f"{name}_west.h5"
    refshim.register_we_file(fname, its, bin_mapper=mapper)
    msm_we.modelWE.processCoordinates = processCoordinates
    model = msm_we.modelWE()
    model.initialize([fname], {"coords": None, "nAtoms": n_atoms, "coord_ndim": coord_ndim}, name,
                     basis_pcoord_bounds=basis, target_pcoord_bounds=target, dim_reduce_method="none", tau=1.0,
                     pcoord_ndim=pcoord_ndim)
    model.get_iterations()
    model.dimReduce()
    clusters = StratifiedClusters(mapper, model, K, [])
    for b, c in enumerate(centres_per_bin):
        if c is not None:
            clusters.cluster_models[b].cluster_centers_ = np.ascontiguousarray(c, dtype=np.float64)
            # what unpickling a fitted MiniBatchKMeans restores and predict() needs
            clusters.cluster_models[b]._n_threads = 1
            clusters.cluster_models[b].n_features_in_ = c.shape[1]
            clusters.cluster_models[b]._n_features_out = c.shape[0]
    if we_remap is not None:
        clusters.we_remap.update(we_remap)
    model.clusters = clusters
    model.n_clusters = K * mapper.nbins
    return model


def _predict_both(model, its_range):
    """Reference StratifiedClusters.predict on parents (processing_from=True) and children (False) per iteration,
    exactly the calls of do_stratified_ray_discretization (_clustering.py:1278-1316)."""
    parents, children = [], []
    for it in its_range:
        model.load_iter_data(it)
        model.get_transition_data_lag0()
        xp = model.coordinates.transform(model.processCoordinates(model.coordPairList[..., 0]))
        xc = model.coordinates.transform(model.processCoordinates(model.coordPairList[..., 1]))
        model.clusters.processing_from = True
        parents.append(model.clusters.predict(xp))
        model.clusters.processing_from = False
        children.append(model.clusters.predict(xc))
    return parents, children


def run_predict_cfg2(n_iters=24):
    """BASELINE config 2 shape (1000 segs x 64-dim, 30 bins x 20 clusters), first ``n_iters`` iterations of the
    seeded host generator, fixed centres: labels from the reference's per-segment predict loop, per-iteration flux
    matrices and the serial get_fluxMatrix."""
    import dataclasses

    import workloads

    msm_we = refshim.load_reference()
    _quiet()
    cfg = dataclasses.replace(workloads.CONFIGS["cfg2"], n_iters=n_iters)
    means, centers = workloads.make_centers(cfg)
    host = workloads.generate_host(cfg, means)
    its = [dict(weights=d["weights"], pcoord=np.stack([d["pcoord0"], d["pcoord1"]], axis=1),
                coords=np.stack([d["parent"], d["child"]], axis=1)[:, :, :, None]) for d in host]
    basis, target = workloads.region_bounds(cfg)
    mapper = refshim.RectilinearBinMapper(workloads.boundaries(cfg))
    model = _model_with_centres(msm_we, "cfg2", its, mapper, centers, cfg.k_per_bin, basis, target, cfg.dim, 1, 1)
    model.launch_ray_discretization()            # the reference's own fan-out + gather
    out = dict(n_iters=np.int64(n_iters),
               input_checksum=np.array(FD.checksum(*[d[k] for d in host for k in ("pcoord0", "pcoord1", "weights", "parent", "child")])),
               centers_checksum=np.array(FD.checksum(*centers)))
    out["dtrajs"] = _cat(model.dtrajs, np.int64).astype(np.int32)
    out["pair_dtrajs"] = np.concatenate([np.asarray(p, dtype=np.int64).reshape(-1, 2) for p in model.pair_dtrajs]).astype(np.int32)
    out["target_bins"] = np.array(sorted(int(b) for b in model.clusters.target_bins), dtype=np.int64)
    out["basis_bins"] = np.array(sorted(int(b) for b in model.clusters.basis_bins), dtype=np.int64)
    model.get_fluxMatrix(0, first_iter=1, last_iter=model.maxIter, use_ray=False)
    r, c = np.nonzero(model.fluxMatrixRaw)
    out.update(flux_row=r.astype(np.int32), flux_col=c.astype(np.int32), flux_val=model.fluxMatrixRaw[r, c],
               flux_shape=np.array(model.fluxMatrixRaw.shape))
    f = model.get_iter_fluxMatrix(5)
    r, c = np.nonzero(f)
    out.update(iter5_row=r.astype(np.int32), iter5_col=c.astype(np.int32), iter5_val=f[r, c])
    path = os.path.join(HERE, "ref_predict_cfg2.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(out['dtrajs'])} frames")


def run_predict_ntl9():
    """The NTL9 fixture's own centres / bin boundaries / we_remap (clustered.obj, via ntl9_clustered.npz) with seeded
    13-D points scattered around them: bin 11 has no centres and is remapped, points fall in basis and target."""
    msm_we = refshim.load_reference()
    _quiet()
    g = np.load(os.path.join(HERE, "ntl9_clustered.npz"))
    nb = len(g["fitted"])
    centres = [g[f"centers_{b}"] if g["fitted"][b] else None for b in range(nb)]
    bnds = [np.asarray(g["boundaries"], dtype=np.float32)]
    mapper = refshim.RectilinearBinMapper(bnds)
    rng = np.random.default_rng(1311)
    lo, hi = float(bnds[0][0]), float(bnds[0][-2]) + 0.06
    D = centres[0].shape[1]
    its = []
    for i in range(6):
        S = 400
        pc = np.stack([rng.uniform(lo, hi, size=S), rng.uniform(lo, hi, size=S)], axis=1)[:, :, None]
        x = np.empty((S, 2, D))
        for t in range(2):
            b = np.array([int(g["we_remap"][int(bb)]) for bb in mapper.assign(pc[:, t])])
            k = rng.integers(0, 25, size=S)
            x[:, t] = np.stack([centres[bb][kk] for bb, kk in zip(b, k)]) + rng.normal(0, 0.05, size=(S, D))
        w = rng.dirichlet(np.ones(S))
        its.append(dict(weights=w, pcoord=pc, coords=x[:, :, :, None]))
    remap = {b: int(g["we_remap"][b]) for b in range(nb)}
    model = _model_with_centres(msm_we, "ntl9", its, mapper, centres, 25, g["basis_bounds"], g["target_bounds"], D, 1, 1,
                                we_remap=remap)
    parents, children = _predict_both(model, range(1, model.maxIter))
    out = FD.pack_iterations([dict(it, parent_id=np.arange(len(it["weights"]))) for it in its])
    out.update(parents=_cat(parents, np.int64), children=_cat(children, np.int64),
               target_bins=np.array(sorted(int(b) for b in model.clusters.target_bins), dtype=np.int64),
               basis_bins=np.array(sorted(int(b) for b in model.clusters.basis_bins), dtype=np.int64))
    # toggle semantics: two predict calls alternate pcoord0List / pcoord1List (stratified_clustering.py:205-210)
    model.load_iter_data(2)
    model.get_transition_data_lag0()
    xp = model.processCoordinates(model.coordPairList[..., 0])
    xc = model.processCoordinates(model.coordPairList[..., 1])
    model.clusters.toggle = True
    model.clusters.processing_from = True
    out["toggle_first"] = model.clusters.predict(xp)
    out["toggle_second"] = model.clusters.predict(xc)
    out["toggle_state_after"] = np.bool_(model.clusters.processing_from)
    path = os.path.join(HERE, "ref_predict_ntl9.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB; basis bins {out['basis_bins']}, target bins {out['target_bins']}")


def run_colour():
    """History-coloured count matrix of WE lineages with the reference's NonMarkovModel.fit (nmm.py:117-167): one
    discrete trajectory per surviving walker of the last iteration, traced back along ``parent_id``; states A / B are
    the basis / target labels."""
    msm_we = refshim.load_reference()
    _quiet()
    from msm_we.nmm import NonMarkovModel

    rng = np.random.default_rng(99)
    n_states, n_iters, S = 12, 30, 40
    labels = [rng.integers(0, n_states, size=S)]
    parents = [np.arange(S)]
    for _ in range(1, n_iters):
        par = rng.integers(0, S, size=S)
        step = rng.integers(-2, 3, size=S)
        labels.append(np.clip(labels[-1][par] + step, 0, n_states - 1))
        parents.append(par)
    trajs = []
    for s in range(S):
        t, cur = [], s
        for it in range(n_iters - 1, -1, -1):
            t.append(int(labels[it][cur]))
            cur = int(parents[it][cur])
        trajs.append(t[::-1])
    stateA, stateB = [0], [n_states - 1]
    nm = NonMarkovModel(trajs, stateA, stateB, lag_time=1, clean_traj=True, sliding_window=True)   # (clean_traj=False renumbers states with a counter that restarts per trajectory)
    out = dict(labels=np.array(labels, dtype=np.int64), parents=np.array(parents, dtype=np.int64),
               n_states=np.int64(n_states), stateA=np.array(stateA), stateB=np.array(stateB),
               nm_cmatrix=np.asarray(nm.nm_cmatrix, dtype=np.float64))
    path = os.path.join(HERE, "ref_colour_lineages.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB, counts {out['nm_cmatrix'].sum():.0f}")


def run_feeder():
    """The reference's own loader on a run spread over TWO west.h5 files (segments of an iteration concatenated in file
    order, _data.py:807-993), a 2-D pcoord of which only the first dimension is loaded, ragged segment counts: what
    ``load_iter_data`` / ``get_iter_coordinates`` / ``get_transition_data_lag0`` leave on the model."""
    msm_we = refshim.load_reference()
    _quiet()
    its = FD.we_dataset(seed=31, n_iters=6, segs0=37, seg_growth=5, n_atoms=3, coord_ndim=3, bins_per_dim=4, k_true=2, pcoord_ndim=2)
    cut = [int(len(d["weights"]) * 0.6) for d in its]
    part = lambda d, sl: {k: v[sl] for k, v in d.items()}          # noqa: E731
    refshim.register_we_file("feeder_a_west.h5", [part(d, slice(0, c)) for d, c in zip(its, cut)])
    refshim.register_we_file("feeder_b_west.h5", [part(d, slice(c, None)) for d, c in zip(its, cut)])
    msm_we.modelWE.processCoordinates = processCoordinates
    model = msm_we.modelWE()
    model.initialize(["feeder_a_west.h5", "feeder_b_west.h5"], {"coords": None, "nAtoms": 3, "coord_ndim": 3}, "feeder",
                     basis_pcoord_bounds=[[0.0, 0.5]], target_pcoord_bounds=[[3.5, 1.0e6]], dim_reduce_method="none",
                     tau=1.0, pcoord_ndim=1)
    model.get_iterations()
    out = FD.pack_iterations(its)
    out.update(cut=np.array(cut, dtype=np.int64), maxIter=np.int64(model.maxIter),
               numSegments=np.asarray(model.numSegments, dtype=np.float64))
    lens, w, p0, p1, west, segind, child, pairs = [], [], [], [], [], [], [], []
    for n in range(1, model.maxIter):
        model.load_iter_data(n)
        lens.append(model.nSeg)
        w.append(np.asarray(model.weightList, dtype=np.float64))
        p0.append(np.asarray(model.pcoord0List, dtype=np.float64).reshape(model.nSeg, -1))
        p1.append(np.asarray(model.pcoord1List, dtype=np.float64).reshape(model.nSeg, -1))
        west.append(np.asarray(model.westList, dtype=np.int64))
        segind.append(np.asarray(model.segindList, dtype=np.int64))
        child.append(np.asarray(model.get_iter_coordinates(n), dtype=np.float64))
        model.get_transition_data_lag0()
        pairs.append(np.asarray(model.coordPairList, dtype=np.float64))
    out.update(f_lens=np.array(lens, dtype=np.int64), f_weights=_cat(w), f_pcoord0=np.concatenate(p0), f_pcoord1=np.concatenate(p1),
               f_west=_cat(west, np.int64), f_segind=_cat(segind, np.int64), f_child=np.concatenate(child),
               f_pairs=np.concatenate(pairs))
    path = os.path.join(HERE, "ref_feeder_twofiles.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB; maxIter {model.maxIter}, segments {lens}")


def main(which):
    if "pipeline1d" in which:
        its = FD.we_dataset(seed=11, n_iters=22, segs0=180, seg_growth=4, n_atoms=5, coord_ndim=3, bins_per_dim=8, k_true=3,
                            skip_bin=5, skip_until=12, regions_last=([[0.0, 0.5]], [[7.5, 1.0e6]]))
        run_pipeline("pipeline1d", its, FD.boundaries(8), [[0.0, 0.5]], [[7.5, 1.0e6]], K=4, pcoord_ndim=1, n_atoms=5,
                     coord_ndim=3, cluster_kwargs={"random_state": 1337},
                     cluster_call_kwargs={"iters_to_use": list(range(1, 11))})
    if "pipeline2d" in which:
        its = FD.we_dataset(seed=12, n_iters=14, segs0=300, seg_growth=0, n_atoms=4, coord_ndim=3, bins_per_dim=4, k_true=2,
                            pcoord_ndim=2)
        run_pipeline("pipeline2d", its, FD.boundaries(4, 2), [[0.0, 0.6], [0.0, 0.6]], [[3.4, 1.0e6], [3.4, 1.0e6]], K=3,
                     pcoord_ndim=2, n_atoms=4, coord_ndim=3, dim_reduce="pca", use_weights=True,
                     cluster_kwargs={"random_state": 7, "init": "random"}, user_mapper=True, organize=False)
    if "pipeline_voronoi" in which:
        # WESTPA's other supported mapper: nearest-centre bins with a user distance function; the cell of centre 5.5 is
        # never seen while clustering (remapped through find_nearest_bin's Voronoi branch, _clustering.py:1331-1396)
        its = FD.we_dataset(seed=13, n_iters=20, segs0=200, seg_growth=3, n_atoms=4, coord_ndim=3, bins_per_dim=8, k_true=3,
                            skip_bin=5, skip_until=11, regions_last=([[0.0, 0.5]], [[7.5, 1.0e6]]))
        run_pipeline("pipeline_voronoi", its, FD.boundaries(8), [[0.0, 0.5]], [[7.5, 1.0e6]], K=4, pcoord_ndim=1, n_atoms=4,
                     coord_ndim=3, cluster_kwargs={"random_state": 4242}, cluster_call_kwargs={"iters_to_use": list(range(1, 10))},
                     voronoi_centers=[[0.5], [1.5], [2.4], [3.6], [4.5], [5.5], [6.5], [7.5]])
    if "feeder" in which:
        run_feeder()
    if "predict_cfg2" in which:
        run_predict_cfg2()
    if "predict_ntl9" in which:
        run_predict_ntl9()
    if "colour" in which:
        run_colour()


if __name__ == "__main__":
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    main(sys.argv[1:] or ["pipeline1d", "pipeline2d", "pipeline_voronoi", "feeder", "predict_cfg2", "predict_ntl9", "colour"])
