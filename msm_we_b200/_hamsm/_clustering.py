"""Clustering + discretization mixin: the reference's method names and results, GPU arithmetic.

reference: msm_we/_hamsm/_clustering.py -- ``cluster_coordinates`` (:142-195), ``cluster_stratified``
(:525-746), ``do_stratified_clustering`` (:748-918), ``organize_stratified`` (:920-1142),
``launch_ray_discretization`` (:1144-1242), ``do_stratified_ray_discretization`` (:1244-1329),
``find_nearest_bin`` (:1331-1396), ``update_cluster_structures`` (:1398-1526), ``get_cluster_centers``
(:1528-1599), ``update_sorted_cluster_centers`` (:1601-1611).

What changed underneath:
* no Ray, no fork-per-iteration ``ProcessPoolExecutor``: ``use_ray`` is accepted and ignored;
* ``do_stratified_clustering`` first PLANS a batch from progress coordinates alone (which iterations it pulls,
  which rows go to which WE bin -- the reference's rule, including its row-pairing quirk), then gathers the
  coordinates once and runs all WE bins of the batch through one K1 + one K2 launch;
* ``launch_ray_discretization`` stages many iterations into one pinned buffer (HDF5 sources read straight into
  it), and K0 + K1 label every parent and child frame of the chunk in a single launch sequence instead of one
  sklearn call per segment;
* ``get_cluster_centers`` / ``update_cluster_structures`` are a device group-by-label (stable sort + per-label
  reduction) instead of one ``np.where`` over every iteration per cluster.
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np

from .._logging import log, ProgressBar
from ..binning import SUPPORTED_MAPPERS as _NATIVE_MAPPERS, RectilinearBinMapper, VoronoiBinMapper
from ..stratified_clustering import StratifiedClusters

# user-extensible, as in the reference (_clustering.py:22)
SUPPORTED_MAPPERS = set(_NATIVE_MAPPERS)

DEFAULT_CHUNK_BYTES = 64 << 20
DEFAULT_RESIDENT_BYTES = 8 << 30      # child feature rows lloyd_refine_clusters may leave on the device for the next pass


class _ResidentChildRows:
    """Feature rows of end-of-segment frames that ``lloyd_refine_clusters`` already shipped to the device, kept for the
    ``launch_ray_discretization`` that follows it (the reference's order: cluster, then discretize the same frames,
    _clustering.py:142-226 -> :1144) so those frames cross PCIe once.  Single use: the discretization pass drops it."""

    def __init__(self, model, X, rows):
        self.X, self.rows = X, rows
        self.key = self.key_of(model)

    @staticmethod
    def key_of(model):
        feat = model.processCoordinates
        return (id(model.iteration_source), id(model.coordinates), getattr(feat, "__func__", feat))

    def lookup(self, model, it, s, D):
        if self.key != self.key_of(model) or self.X.shape[1] != D:
            return None
        hit = self.rows.get(it)
        return self.X[hit[0]:hit[0] + s] if hit is not None and hit[1] == s else None

    # a copy or a pickle of the model does not take the device rows along
    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_nothing, ())


def _nothing():
    return None


_STAGING_POOL = None


def _staging_pool():
    """Threads that fill the pinned staging buffers.  One core copies pageable -> pinned memory at ~6 GB/s,
    a ninth of what the PCIe link then moves, so the per-iteration featurise + copy jobs are spread over a
    few threads (numpy releases the GIL inside large copies).  ``MSM_WE_B200_STAGING_THREADS=1`` disables."""
    global _STAGING_POOL
    if _STAGING_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor

        n = int(os.environ.get("MSM_WE_B200_STAGING_THREADS", "0")) or min(8, max(1, (os.cpu_count() or 2) // 2))
        _STAGING_POOL = (ThreadPoolExecutor(max_workers=n, thread_name_prefix="mwe-stage") if n > 1 else False)
    return _STAGING_POOL or None


def connected_sets(C, directed=True):
    """Strongly connected components of the graph whose edges are the non-zero entries of ``C``, each as a sorted
    index array, largest first (components of equal size in order of their first discovery label) -- the contract
    of the reference's ``find_connected_sets`` (msm_we/utils.py:21-84, itself msmtools' routine); host-side graph
    work on an (n+2)^2 matrix, as in the reference."""
    from scipy.sparse import csr_matrix, issparse
    from scipy.sparse.csgraph import connected_components

    if not issparse(C):
        C = csr_matrix(C)
    nc, comp = connected_components(C, directed=directed, connection="strong")
    order = np.argsort(comp, kind="stable")
    bounds = np.concatenate([[0], np.cumsum(np.bincount(comp, minlength=nc))])
    sets = [np.sort(order[bounds[i]:bounds[i + 1]]) for i in range(nc)]
    return sorted(sets, key=lambda s: -len(s))


class _RemoteShim:
    """``obj.method.remote(...)`` compatibility for code written against the ``@ray.remote`` functions:
    runs synchronously and returns the result itself."""

    def __init__(self, fn):
        self._fn = fn
        self.__doc__ = fn.__doc__

    def __get__(self, obj, objtype=None):
        return self

    def __call__(self, *args, **kwargs):
        return self._fn(*args, **kwargs)

    def remote(self, *args, **kwargs):
        return self._fn(*args, **kwargs)


def _do_stratified_ray_discretization(model, kmeans_model, iteration, processCoordinates):
    """One iteration: load pcoords / weights / coordinate pairs, featurise + transform parents and
    children, ``predict`` twice (reference :1244-1329).  Empty iterations return the reference's
    3-tuple ``(None, 0, iteration)``."""
    self = model
    # the reference's caller detaches .model before shipping the clusterer (:1171-1174); copying it with the model
    # attached would duplicate everything the model holds
    original, attached = kmeans_model, kmeans_model.model
    original.model = None
    try:
        kmeans_model = deepcopy(original)
    finally:
        original.model = attached
    kmeans_model.model = self
    self.load_iter_data(iteration)
    self.get_transition_data_lag0()
    parent_coords, child_coords = self.coordPairList[..., 0], self.coordPairList[..., 1]
    if child_coords.shape[0] == 0:
        return None, 0, iteration
    transformed_parent = self.coordinates.transform(processCoordinates(parent_coords))
    transformed_child = self.coordinates.transform(processCoordinates(child_coords))
    try:
        kmeans_model.processing_from = True
        parent_dtrajs = kmeans_model.predict(transformed_parent)
        kmeans_model.processing_from = False
        child_dtrajs = kmeans_model.predict(transformed_child)
    except AttributeError as e:
        log.error("Cluster center was not initialized and not remapped")
        log.error(kmeans_model.we_remap)
        raise e
    return (parent_dtrajs, child_dtrajs), 1, iteration, kmeans_model.target_bins, kmeans_model.basis_bins


class ClusteringMixin:
    n_clusters = None
    clusters = None
    clusterFile = None
    use_weights_in_clustering = False
    targetRMSD_centers = None
    targetRMSD_minmax = None
    targetRMSD_all = None
    all_centers = None
    sorted_centers = None
    pcoord_cache = None
    cluster_structures = None
    cluster_structure_weights = None
    pre_discretization_model = None
    post_cluster_model = None
    _resident_child_rows = None

    do_stratified_ray_discretization = _RemoteShim(_do_stratified_ray_discretization)

    @staticmethod
    def check_connect_ray():
        """The GPU path has no Ray cluster to connect to (reference: msm_we.py:446-460)."""
        return None

    # ------------------------------------------------------------------------------------------
    def cluster_coordinates(self, n_clusters, streaming=False, first_cluster_iter=None, use_ray=False, stratified=True,
                            iters_to_use=None, store_validation_model=False, progress_bar=None, **_cluster_args):
        """reference: _clustering.py:142-195."""
        self.clustering_method = None
        log.info("Be aware: Number of cluster centers is an important parameter, and can drastically affect model "
                 "quality. We recommend examining block-validation results with a range of numbers of clusters, to "
                 "check for overfitting.")
        if stratified:
            self.clustering_method = "stratified"
            self.cluster_stratified(n_clusters=n_clusters, streaming=streaming, first_cluster_iter=first_cluster_iter,
                                    use_ray=use_ray, iters_to_use=iters_to_use, progress_bar=progress_bar,
                                    **_cluster_args)
        else:
            self.clustering_method = "aggregated"
            self.cluster_aggregated(n_clusters=n_clusters, streaming=streaming, first_cluster_iter=first_cluster_iter,
                                    use_ray=use_ray, iters_to_use=iters_to_use, **_cluster_args)
        if store_validation_model:
            self.post_cluster_model = deepcopy(self)

    def cluster_aggregated(self, *args, **kwargs):
        # Appendix A.13 of SURVEY.md: the flux path reads pair_dtrajs, which only the stratified
        # discretization writes; aggregate clustering is outside the hot path this package covers.
        raise NotImplementedError("msm_we_b200 implements the stratified clustering path only "
                                  "(reference: _clustering.py:197-523 is out of scope)")

    # ------------------------------------------------------------------------------------------
    def cluster_stratified(self, n_clusters, streaming=True, first_cluster_iter=None, use_ray=True, bin_iteration=2,
                           iters_to_use=None, user_bin_mapper=None, progress_bar=None, **_cluster_args):
        """reference: _clustering.py:525-746."""
        if user_bin_mapper is not None:
            log.info("Loading user-specified bin mapper for stratified clustering.")
            bin_mapper = user_bin_mapper
        else:
            bin_mapper = self._load_bin_mapper(bin_iteration)
            if type(bin_mapper) not in SUPPORTED_MAPPERS:
                log.warning(f"{type(bin_mapper)} mapper loaded, but supported mappers are {SUPPORTED_MAPPERS} and "
                            f"others may produce inconsistent bins between iterations. Please provide a supported "
                            f"user_bin_mapper.")
                raise Exception

        ignored_bins = []
        if not streaming or not use_ray:
            log.debug("Stratified clustering always streams; use_ray is ignored on the GPU path.")
            streaming = True
            use_ray = True

        stratified_clusters = StratifiedClusters(bin_mapper, self, n_clusters, ignored_bins, **_cluster_args)

        if iters_to_use is None and first_cluster_iter is None:
            first_cluster_iter = 1
            iters_to_use = range(first_cluster_iter, self.maxIter)
        elif iters_to_use is None and first_cluster_iter is not None:
            iters_to_use = range(first_cluster_iter, self.maxIter)
        elif iters_to_use is not None and first_cluster_iter is not None:
            log.error("Conflicting parameters -- either iters_to_use OR first_cluster_iter should be provided, not both.")

        self.dtrajs = []
        extra_iters_used = 0
        all_filled_bins = set()
        all_unfilled_bins = set()

        with ProgressBar(progress_bar) as progress:
            task = progress.add_task(description="Clustering", total=len(iters_to_use), completed=0)
            for iter_idx, iteration in enumerate(iters_to_use):
                if extra_iters_used > 0:
                    extra_iters_used -= 1
                    log.debug(f"Already processed  iter  {iteration}")
                    continue
                ignored_bins = []
                filled_bins, unfilled_bins = [], []
                try:
                    stratified_clusters, extra_iters_used, filled_bins, unfilled_bins = self.do_stratified_clustering(
                        [self, stratified_clusters, iters_to_use[iter_idx:], self.processCoordinates, ignored_bins])
                except AssertionError as e:
                    if iter_idx == 0:
                        log.info(f"Failed with {iter_idx} + {extra_iters_used} vs len {(len(iters_to_use))}")
                        raise e
                    log.info("Clustering couldn't use last iteration, not all bins filled.")
                all_filled_bins.update(int(b) for b in filled_bins)
                all_unfilled_bins.update(int(b) for b in unfilled_bins)
                progress.update(task, advance=1 + extra_iters_used)

        true_unfilled = np.setdiff1d(range(bin_mapper.nbins), list(all_filled_bins))
        for unfilled_bin_idx in true_unfilled:
            remap_bin = self.find_nearest_bin(bin_mapper, unfilled_bin_idx, list(all_filled_bins))
            stratified_clusters.we_remap[int(unfilled_bin_idx)] = int(remap_bin)
            log.debug(f"Remapped {unfilled_bin_idx} to {remap_bin}")

        self.clusters = stratified_clusters
        self.clusters.model = self
        self.n_clusters = n_clusters * (bin_mapper.nbins)
        self.clusters.toggle = False
        self.launch_ray_discretization(progress_bar)

    def _load_bin_mapper(self, bin_iteration=2):
        """The reference unpickles the mapper WESTPA stored in the HDF5 file (``analysis.Run(file).iteration(n)
        .bin_mapper``, :588-590).  That needs westpa; a model can also carry one in ``self.bin_mapper``."""
        bin_mapper = getattr(self, "bin_mapper", None)
        if bin_mapper is not None:
            return bin_mapper
        try:
            from westpa import analysis
        except ImportError:
            raise Exception("No bin mapper: the reference unpickles it from the WESTPA HDF5 file with westpa, which is "
                            "not importable here; pass user_bin_mapper= or set model.bin_mapper") from None
        log.debug(f"Obtaining bin definitions from iteration {bin_iteration} in file {self.fileList[0]}")
        return analysis.Run(self.fileList[0]).iteration(bin_iteration).bin_mapper

    # ------------------------------------------------------------------------------------------
    def _plan_stratified_batch(self, bin_mapper, min_coords, iters_to_use):
        """The batching rule of do_stratified_clustering (:794-886), evaluated on progress coordinates alone.

        Iterations are pulled until every WE bin seen so far holds >= ``min_coords`` usable segments (parents outside
        basis / target); if the iterations run out, under-filled bins are folded into the nearest filled bin.  Returns
        ``(pulled, rows_per_iter, assignments, unique_bins, unfilled_bins)``: ``assignments[j]`` is the WE bin of the
        j-th USABLE segment of the pulled iterations.  (The reference then uses j as a row number of the UNFILTERED
        coordinate array, :892-899 / SURVEY Appendix A.5; the caller reproduces that.)"""
        P = self.pcoord_ndim
        pulled, rows_per_iter, kept_bins = [], [], []
        counts = np.zeros(bin_mapper.nbins, dtype=np.int64)
        unfilled_bins = []
        pos = 0
        while True:
            if pos >= len(iters_to_use):
                seen = np.flatnonzero(counts)
                log.warning(f"At iteration {pulled[-1] if pulled else None} (pulled {len(pulled) - 1} extra), couldn't get "
                            f"segments in all bins, and no iterations left.")
                unfilled_bins = seen[counts[seen] < min_coords]
                filled_bins = np.setdiff1d(seen, unfilled_bins)
                assignments = np.concatenate(kept_bins) if kept_bins else np.array([])
                for unfilled_bin in unfilled_bins:
                    nearest = self.find_nearest_bin(bin_mapper, unfilled_bin, list(filled_bins))
                    log.warning(f"Remapping segments from unfilled bin {unfilled_bin} to {nearest} for stratified clustering")
                    assignments[assignments == unfilled_bin] = nearest
                return pulled, rows_per_iter, assignments, filled_bins, unfilled_bins
            iteration = iters_to_use[pos]
            if iteration > self.maxIter:
                log.warning(f"At iteration {iteration} (pulled {len(pulled) - 1} extra), couldn't get segments in all "
                            f"bins, and no iterations left")
                break
            pos += 1
            if self.iteration_source.has(iteration):
                rec = self._record(iteration, coords=False)
                pc0 = rec.pcoord0[:, :P]
            else:
                pc0 = np.empty((0, P))
            pulled.append(iteration)
            rows_per_iter.append(pc0.shape[0])
            usable = ~(self.is_WE_target(pc0) | self.is_WE_basis(pc0)) if pc0.shape[0] else np.zeros(0, dtype=bool)
            if usable.any():
                b = np.asarray(bin_mapper.assign(pc0[usable])).astype(np.int64)
                kept_bins.append(b)
                counts += np.bincount(b, minlength=bin_mapper.nbins)
            seen = np.flatnonzero(counts)
            # np.all of an empty comparison is True: an iteration whose parents are all in the basis "fills" nothing
            if np.all(counts[seen] >= min_coords):
                break
        assignments = np.concatenate(kept_bins) if kept_bins else np.array([])
        return pulled, rows_per_iter, assignments, np.flatnonzero(counts), unfilled_bins

    def do_stratified_clustering(self, arg):
        """reference: _clustering.py:748-918.  Returns ``(kmeans_models, used_iters, unique_bins,
        unfilled_bins)``.  The per-bin ``partial_fit`` calls of the reference (:890-916) are issued as ONE
        batched GPU step over all bins of the batch."""
        from ..clustering_ops import partial_fit_models

        self, kmeans_models, iters_to_use, processCoordinates, ignored_bins = arg
        bin_mapper = kmeans_models.bin_mapper
        min_coords = kmeans_models.cluster_args["n_clusters"]
        pulled, rows_per_iter, assignments, unique_bins, unfilled_bins = self._plan_stratified_batch(
            bin_mapper, min_coords, list(iters_to_use))
        used_iters = len(pulled) - 1

        # Row j of the FILTERED assignment array indexes row j of the UNFILTERED coordinates / weights (the reference's
        # pairing, :892-899), so only the first len(assignments) unfiltered rows are ever touched.
        need = int(assignments.shape[0])
        coords, weights, have = [], [], 0
        for k, iteration in enumerate(pulled):
            c = self.get_iter_coordinates(iteration)           # drops rows with NaN coordinates (:543-553)
            assert c.shape[0] == rows_per_iter[k], f"({rows_per_iter[k]}, {self.pcoord_ndim}), {c.shape}"
            if have < need:
                coords.append(c)
                # (the reference appends later iterations' weights only when they are used, :856-857)
                weights.append(self.seg_weights[iteration] if (k == 0 or self.use_weights_in_clustering) else None)
                have += c.shape[0]
        iter_coords = np.concatenate(coords, axis=0) if coords else np.empty((0, self.nAtoms or 0, self.coord_ndim or 3))
        seg_weights = np.concatenate([w for w in weights if w is not None]) if self.use_weights_in_clustering and weights \
            else (weights[0] if weights else np.array([]))

        batch = []
        for _bin in unique_bins:
            segs_in_bin = np.argwhere(assignments == _bin)
            transformed_coords = self.coordinates.transform(processCoordinates(np.squeeze(iter_coords[segs_in_bin])))
            w = seg_weights[segs_in_bin].squeeze() if self.use_weights_in_clustering else None
            batch.append((kmeans_models.cluster_models[int(_bin)], transformed_coords, w))
        try:
            partial_fit_models(batch)
        except ValueError as e:
            log.error(f"Error fitting k-means in bins {list(unique_bins)}")
            raise e
        return kmeans_models, used_iters, unique_bins, unfilled_bins

    # ------------------------------------------------------------------------------------------
    def launch_ray_discretization(self, progress_bar=None):
        """reference: _clustering.py:1144-1242.  Sets ``self.dtrajs`` (child labels per iteration) and
        ``self.pair_dtrajs`` (``[S, 2]`` int64 arrays of (parent, child) labels, which is what
        ``np.array(list(zip(parent, child)))`` gives the flux code)."""
        import torch

        from .._pinning import PINS
        from .. import ops as _ops

        self.check_connect_ray()
        self.dtrajs = []
        if self.pre_discretization_model is None:
            # (iteration sources are shared by copies, so this snapshots the model state, not the data set)
            self.pre_discretization_model = deepcopy(self)
        else:
            log.debug("Using cached model for discretization")

        clusters = self.clusters
        dev = clusters.device_state()
        chunk_bytes = int(clusters.cluster_args.get("gpu_chunk_bytes", DEFAULT_CHUNK_BYTES))
        n_iters = self.maxIter - 1
        dtrajs = [None] * n_iters
        pair_dtrajs = [None] * n_iters
        D = dev.D
        P = self.pcoord_ndim
        src = self.iteration_source
        seg_counts = [int(self.numSegments[it - 1]) if self.numSegments is not None and it <= len(self.numSegments)
                      else None for it in range(1, self.maxIter)]

        featurise, transform = self.processCoordinates, self.coordinates.transform
        # a fitted linear projection (the reference's PCA ``coordinates.transform``) is applied on the device:
        # the featurised frames are shipped as they are
        projection = getattr(self.coordinates, "device_projection", None)
        projection = projection(dev.device) if callable(projection) else None
        identity_transform = projection is not None or getattr(self.coordinates, "is_identity", False)
        # structures can be read from the source straight into the pinned rows when nothing but a flatten stands
        # between the stored structure and the feature row
        default_featuriser = getattr(getattr(type(self), "processCoordinates", None), "_mwe_flatten", False) \
            and "processCoordinates" not in self.__dict__
        direct_read = default_featuriser and identity_transform and hasattr(src, "read_pair_into")
        Din_hint = None
        if projection is not None:
            Din_hint = int(projection[0].shape[1])
        row_bytes = 2 * ((Din_hint or D) + P) * 8

        # plan chunks of whole iterations; each chunk is staged in ONE pinned buffer (features | pcoords),
        # copied with one async H2D and labelled by one K0 + K1 launch sequence.  Two staging buffers
        # alternate, so the host fills chunk i+1 while the GPU copies and labels chunk i.
        chunks, cur, cur_rows = [], [], 0
        for it in range(1, self.maxIter):
            s = seg_counts[it - 1]
            if s is None:
                self.load_iter_data(it)
                s = self.nSeg
            if s == 0:
                continue
            if cur and (cur_rows + s) * row_bytes > chunk_bytes:
                chunks.append(cur)
                cur, cur_rows = [], 0
            cur.append((it, s))
            cur_rows += s
        if cur:
            chunks.append(cur)

        resident = getattr(self, "_resident_child_rows", None)
        if resident is not None and (projection is not None or resident.key != resident.key_of(self)):
            resident = None

        stream = torch.cuda.current_stream()
        slots = [dict(event=None, host=None, n=0), dict(event=None, host=None, n=0)]
        inflight = []   # (chunk, n, labels_host, bins_host, flags_host, nan_host, done_event)

        def source_owned(feat, rec):
            """True when ``feat`` is a view of an array the iteration source holds for the model's lifetime -- only
            those may be page-locked in place (a featuriser's temporary would be unlocked and freed while its
            asynchronous copy is still queued)."""
            if not getattr(src, "owns_arrays", False):
                return False
            owner = PINS._owner(feat)
            return any(owner is PINS._owner(a) for a in (rec.parent_coords, rec.child_coords) if a is not None)

        def stage(ci, chunk):
            n = sum(s for _, s in chunk)
            slot = slots[ci % 2]
            if slot["event"] is not None:
                slot["event"].synchronize()          # the H2D that last read this buffer has finished
            pool = _staging_pool()

            def featurise_pair(item):
                parent_coords, child_coords = self.iter_coordinate_pair(item[0])
                if resident is not None and resident.lookup(self, item[0], item[1], D) is not None:
                    return np.asarray(transform(featurise(parent_coords))), None     # child rows are on the device already
                fp, fc = featurise(parent_coords), featurise(child_coords)
                if projection is None:
                    fp, fc = transform(fp), transform(fc)
                return np.asarray(fp), np.asarray(fc)

            feats = None
            if direct_read:
                Din = Din_hint or D
            else:
                # user featurisers / host transforms run on the staging threads (numpy releases the GIL in its loops);
                # a featuriser that only reshapes is cheaper inline than a hand-off to a thread
                import time as _time

                t0 = _time.perf_counter()
                feats = [featurise_pair(chunk[0])]
                if pool is not None and len(chunk) > 1 and _time.perf_counter() - t0 > 2e-4:
                    feats += list(pool.map(featurise_pair, chunk[1:]))
                else:
                    feats += [featurise_pair(c) for c in chunk[1:]]
                Din = feats[0][0].shape[1] if projection is not None else D
            # pinned staging: pcoords always (small), feature rows only for iterations whose arrays cannot be
            # page-locked in place
            need = 2 * n * (Din + P)
            if slot["host"] is None or slot["host"].numel() < need:
                slot["host"] = torch.empty(need, dtype=torch.float64, pin_memory=True)
            host = slot["host"]
            hx_t = host[: 2 * n * Din].view(2 * n, Din)
            hp_t = host[2 * n * Din: need].view(2 * n, P)
            hx, hp = hx_t.numpy(), hp_t.numpy()
            X = torch.empty((2 * n, Din), dtype=torch.float64, device=dev.device)

            pos, jobs, staged = 0, [], []
            for k, (it, s) in enumerate(chunk):
                rec = self._record(it, coords=not direct_read)
                hp[pos:pos + s] = rec.pcoord0[:, :P]
                hp[n + pos:n + pos + s] = rec.pcoord1[:, :P]
                if direct_read:
                    def read(it=it, pos=pos, s=s):
                        if not src.read_pair_into(it, hx[pos:pos + s], hx[n + pos:n + pos + s]):
                            pc_, cc_ = self.iter_coordinate_pair(it)
                            np.copyto(hx[pos:pos + s], pc_.reshape(s, -1))
                            np.copyto(hx[n + pos:n + pos + s], cc_.reshape(s, -1))
                    if pool is None:
                        read()
                    else:
                        jobs.append(pool.submit(read))
                    staged += [(pos, s), (n + pos, s)]
                    pos += s
                    continue
                fp, fc = feats[k]
                if fc is None:
                    X[n + pos:n + pos + s].copy_(resident.lookup(self, it, s, D))
                for off, feat in ((pos, fp), (n + pos, fc)):
                    if feat is None:
                        continue
                    if feat.shape != (s, Din):
                        raise ValueError(f"featurised coordinates of iteration {it} have shape {feat.shape}, expected {(s, Din)}")
                    if source_owned(feat, rec) and PINS.ensure(feat):
                        # the model's own array is page-locked: straight to the device at link speed
                        X[off:off + s].copy_(torch.from_numpy(feat), non_blocking=True)
                    else:
                        # featurised / projected on the fly (or not lockable): stage through the pinned rows
                        if pool is None:
                            np.copyto(hx[off:off + s], feat)
                        else:
                            jobs.append(pool.submit(np.copyto, hx[off:off + s], feat))
                        staged.append((off, s))
                pos += s
            for j in jobs:
                j.result()
            staged.sort()
            k = 0
            while k < len(staged):                   # one copy per run of adjacent staged row ranges
                lo, hi = staged[k][0], staged[k][0] + staged[k][1]
                while k + 1 < len(staged) and staged[k + 1][0] == hi:
                    k += 1
                    hi = staged[k][0] + staged[k][1]
                X[lo:hi].copy_(hx_t[lo:hi], non_blocking=True)
                k += 1
            Xin = X
            if projection is not None:
                X = _ops.project(X, projection[0], projection[1])
            Pc = hp_t.to(dev.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            slot["event"] = ev
            labels, bins, flags = dev.predict(X, Pc, pcoord_host=hp)
            # segments with a NaN start / end coordinate (their transition weight is zeroed by the flux pass,
            # _data.py:302-313): found here, where the coordinates already are on the device
            nan_rows = _ops.rows_with_nan(Xin) if direct_read else None
            lh = torch.empty(2 * n, dtype=torch.int64, pin_memory=True)
            bh = torch.empty(2 * n, dtype=torch.int32, pin_memory=True)
            fh = torch.empty(2 * n, dtype=torch.uint8, pin_memory=True)
            nh = None
            lh.copy_(labels, non_blocking=True); bh.copy_(bins, non_blocking=True); fh.copy_(flags, non_blocking=True)
            if nan_rows is not None:
                nh = torch.empty(2 * n, dtype=torch.uint8, pin_memory=True)
                nh.copy_(nan_rows, non_blocking=True)
            done = torch.cuda.Event()
            done.record(stream)
            inflight.append((chunk, n, lh, bh, fh, nh, done))

        with ProgressBar(progress_bar) as progress:
            task = progress.add_task(description="Discretizing trajectories", total=n_iters)

            def collect(item):
                chunk, n, lh, bh, fh, nh, done = item
                done.synchronize()
                labels_h, bins_h, flags_h = lh.numpy(), bh.numpy(), fh.numpy()
                is_target = (flags_h & 2) != 0
                is_basis = ((flags_h & 1) != 0) & ~is_target
                clusters.target_bins.update(int(b) for b in np.unique(bins_h[is_target]))
                clusters.basis_bins.update(int(b) for b in np.unique(bins_h[is_basis]))
                # one [n, 2] (parent, child) array and one child array per chunk; the per-iteration entries
                # are row ranges of them
                pairs_all = np.empty((n, 2), dtype=np.int64)
                pairs_all[:, 0] = labels_h[:n]
                pairs_all[:, 1] = labels_h[n:2 * n]
                child_all = labels_h[n:2 * n].copy()
                nan_seg = None
                if nh is not None:
                    nan_h = nh.numpy()
                    nan_seg = (nan_h[:n] | nan_h[n:2 * n]) != 0
                    if not nan_seg.any():
                        nan_seg = False
                pos = 0
                for it, s in chunk:
                    dtrajs[it - 1] = child_all[pos:pos + s]
                    pair_dtrajs[it - 1] = pairs_all[pos:pos + s]
                    if nan_seg is not None:
                        self.note_nan_segments(it, np.zeros(0, np.int64) if nan_seg is False
                                               else np.flatnonzero(nan_seg[pos:pos + s]))
                    pos += s
                progress.update(task, advance=len(chunk))

            # the results of chunk i are unpacked while chunk i+1 is being copied and labelled
            for ci, chunk in enumerate(chunks):
                stage(ci, chunk)
                if ci >= 1:
                    collect(inflight[ci - 1])
            if inflight:
                collect(inflight[-1])
            dev.check_errors()

        self._resident_child_rows = None       # single use: the next pass ships its own frames
        self.dtrajs = [d for d in dtrajs if d is not None]
        self.pair_dtrajs = [d for d in pair_dtrajs if d is not None]
        log.debug("Discretization complete")

    # ------------------------------------------------------------------------------------------
    def lloyd_refine_clusters(self, n_iter=10, iters_to_use=None, max_device_bytes=None):
        """Full-batch Lloyd refinement of every WE bin's cluster model from its current centres -- the stratified
        counterpart of the ``KMeans.fit`` the reference runs on its aggregated path (_clustering.py:289,491; sklearn
        ``lloyd_iter_chunked_dense``), and what BASELINE config 5 calls "10 Lloyd iters".  The end-of-segment frames of
        ``iters_to_use`` (default: every discretizable iteration) are featurised, shipped to the GPU once and kept
        resident for all ``n_iter`` iterations; as in the streaming clustering they are binned by the PARENT progress
        coordinate and frames whose parent sits in the basis / target are left out (:849-877).  Updates
        ``cluster_models[b].cluster_centers_`` in place; returns the number of frames used."""
        import torch

        from .._pinning import PINS
        from .. import ops as _ops
        from ..clustering_ops import lloyd_fit

        clusters = self.clusters
        dev = clusters.device_state()
        iters = list(range(1, self.maxIter) if iters_to_use is None else iters_to_use)
        P = self.pcoord_ndim
        src = self.iteration_source
        counts = [int(src.n_segments(it)) if src.has(it) else 0 for it in iters]
        n = int(sum(counts))
        if n == 0:
            return 0
        projection = getattr(self.coordinates, "device_projection", None)
        projection = projection(dev.device) if callable(projection) else None
        Din = int(projection[0].shape[1]) if projection is not None else dev.D
        need = n * (Din + P + 4) * 8
        if max_device_bytes is not None and need > max_device_bytes:
            raise MemoryError(f"{n} frames x {Din} features ({need} bytes) exceed max_device_bytes={max_device_bytes}")
        try:
            X = torch.empty((n, Din), dtype=torch.float64, device=dev.device)
        except torch.OutOfMemoryError as e:
            raise MemoryError(f"{n} frames x {Din} features do not fit on the device; pass fewer iterations "
                              f"(iters_to_use=) -- config 5 is sized for iteration-range shards over 8 GPUs") from e
        hp_t = torch.empty((n, P), dtype=torch.float64, pin_memory=True)
        hp = hp_t.numpy()
        stage_rows = max(counts)
        stage, stage_ev, n_staged = None, [None, None], 0
        pos = 0
        row_of_iter = {}
        for it, s in zip(iters, counts):
            if s == 0:
                continue
            row_of_iter[it] = (pos, s)
            rec = self._record(it)
            feat = self.processCoordinates(self._as_structures(rec.child_coords))
            if projection is None:
                feat = self.coordinates.transform(feat)
            feat = np.asarray(feat)
            if feat.shape != (s, Din):
                raise ValueError(f"featurised coordinates of iteration {it} have shape {feat.shape}, expected {(s, Din)}")
            hp[pos:pos + s] = rec.pcoord0[:, :P]
            owned = getattr(src, "owns_arrays", False) and PINS._owner(feat) is PINS._owner(rec.child_coords)
            if owned and PINS.ensure(feat):
                X[pos:pos + s].copy_(torch.from_numpy(feat), non_blocking=True)
            else:
                if stage is None:
                    stage = [torch.empty((stage_rows, Din), dtype=torch.float64, pin_memory=True) for _ in range(2)]
                k = n_staged & 1
                n_staged += 1
                if stage_ev[k] is not None:
                    stage_ev[k].synchronize()
                np.copyto(stage[k].numpy()[:s], feat)
                X[pos:pos + s].copy_(stage[k][:s], non_blocking=True)
                stage_ev[k] = torch.cuda.Event()
                stage_ev[k].record(torch.cuda.current_stream())
            pos += s
        if projection is not None:
            X = _ops.project(X, projection[0], projection[1])
        pc = hp_t.to(dev.device, non_blocking=True)
        bins, flags = dev.bins_and_flags(pc, pcoord_host=hp)
        centers = dev.centers.clone()
        lloyd_fit(X, None, bins, centers, dev.bin_offset, dev.max_k, int(n_iter), flags_dev=flags, errors=dev.errors)
        centers_h = centers.cpu().numpy()
        dev.check_errors()
        offs = dev.bin_offset_host
        for b, m in enumerate(clusters.cluster_models):
            if hasattr(m, "cluster_centers_") and offs[b + 1] > offs[b]:
                m.cluster_centers_ = np.ascontiguousarray(centers_h[offs[b]:offs[b + 1]])
        clusters.adopt_device_centers(centers)       # the device snapshot follows without a second upload
        # the discretization that normally follows labels the same end-of-segment frames: leave them on the device
        # for it when they are small enough (features after ``coordinates.transform``; not with a device projection,
        # whose input rows are what the discretization ships)
        budget = int(clusters.cluster_args.get("gpu_resident_bytes", DEFAULT_RESIDENT_BYTES))
        self._resident_child_rows = (_ResidentChildRows(self, X, row_of_iter)
                                     if projection is None and X.numel() * 8 <= budget else None)
        return n

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def find_nearest_bin(bin_mapper, bin_idx, filled_bins):
        """reference: _clustering.py:1331-1396."""
        assert len(filled_bins) > 0, ("Can't find nearest populated bin -- no WE bins are populated with clusters! "
                                      "Try fewer clusters/bin.")
        is_voronoi = isinstance(bin_mapper, VoronoiBinMapper) or type(bin_mapper).__name__ == "VoronoiBinMapper"
        is_rect = isinstance(bin_mapper, RectilinearBinMapper) or type(bin_mapper).__name__ == "RectilinearBinMapper"
        assert is_voronoi or is_rect, f"{type(bin_mapper)} is unsupported!"
        if is_voronoi:
            centers = np.asarray(bin_mapper.centers)
            distance_function = bin_mapper.dfunc
        else:
            def _rmsd(point, _centers):
                return np.sqrt(np.mean(np.power(point - _centers, 2), axis=1))

            distance_function = _rmsd
            _centers = []
            for dim in bin_mapper.boundaries:
                dim = np.asarray(dim)
                _centers.append(dim[:-1] + (dim[1:] - dim[:-1]) / 2)
            centers = np.array(np.meshgrid(*_centers)).T.squeeze().reshape(-1, len(bin_mapper.boundaries))
        all_ignored = np.setdiff1d(range(centers.shape[0]), filled_bins)
        other_centers = np.delete(centers, all_ignored, axis=0)
        closest = np.argmin(distance_function(centers[int(bin_idx)], other_centers))
        for _bin_idx in sorted(all_ignored):
            if closest >= _bin_idx:
                closest += 1
        return closest

    # ------------------------------------------------------------------------------------------
    def organize_stratified(self, use_ray=True, progress_bar=None):
        """reference: _clustering.py:920-1142.  Removes every cluster outside the largest strongly connected set
        of the raw flux matrix from its WE bin's model, remaps WE bins that lost all their clusters, then runs the
        hot path a second time: re-discretize (K0 + K1), per-cluster pcoord statistics (group-by-label), re-flux
        (K3), and orders the cleaned matrix by mean pcoord."""
        fmatrix_original = self.fluxMatrixRaw.copy()
        fmatrix = self.fluxMatrixRaw.copy()
        fmatrix[-1, -2] = 1.0
        sets = connected_sets(fmatrix, directed=True)
        start_cleaning_idx = 1
        if len(sets) == 1:
            log.info("Nothing to clean")
            states_to_remove = np.array([], dtype=np.int64)
        else:
            states_to_remove = np.concatenate(sets[start_cleaning_idx:])
            log.debug(f"Cleaning states {states_to_remove}")

        models = self.clusters.cluster_models
        nbins = self.clusters.bin_mapper.nbins
        sizes = np.array([len(m.cluster_centers_) if hasattr(m, "cluster_centers_") else 0 for m in models], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(sizes)])
        remove = np.zeros(max(int(offsets[-1]), self.fluxMatrixRaw.shape[0]), dtype=bool)
        remove[states_to_remove] = True
        empty_we_bins = set()
        for we_bin in range(nbins):
            lo, hi = int(offsets[we_bin]), int(offsets[we_bin + 1])
            drop = np.flatnonzero(remove[lo:hi])
            if drop.size == 0:
                if hi == lo:
                    empty_we_bins.add(we_bin)
                continue
            if drop.size == hi - lo:
                empty_we_bins.add(we_bin)
            # (a bin that loses everything keeps a zero-row centre array, as np.delete leaves it in the reference)
            models[we_bin].cluster_centers_ = np.delete(models[we_bin].cluster_centers_, drop, 0)

        log.info(f"Started with {self.n_clusters} clusters, and removed {len(states_to_remove)}")
        self.n_clusters = self.n_clusters - len(states_to_remove)
        assert self.n_clusters > 1, "All clusters would be cleaned! You probably need more data, fewer clusters, or both."

        populated_we_bins = np.setdiff1d(range(nbins), list(empty_we_bins))
        if len(empty_we_bins) > 0:
            log.warning(f"All clusters were cleaned from bins {empty_we_bins} (This is normal for the source/target WE "
                        f"bins, and this can be ignored if only those are listed here.)")
        for empty_we_bin in empty_we_bins:
            self.clusters.we_remap[empty_we_bin] = self.find_nearest_bin(self.clusters.bin_mapper, empty_we_bin,
                                                                         populated_we_bins)
        for we_bin in range(nbins):
            target = self.clusters.we_remap[we_bin]
            if not hasattr(models[target], "cluster_centers_"):
                log.error(f"Error obtaining clusters for WE bin {we_bin}, remapped to {target}. "
                          f"Target {self.clusters.target_bins}, basis {self.clusters.basis_bins}")
                raise AttributeError(f"'{type(models[target]).__name__}' object has no attribute 'cluster_centers_'")

        # re-discretize on the cleaned centres
        self.clusters.toggle = False
        self.clusters.processing_from = False
        self.launch_ray_discretization(progress_bar=progress_bar)

        pcoord_sort_indices = self.get_cluster_centers()

        # and recalculate the flux matrix (the toggle is vestigial: get_fluxMatrix reads pair_dtrajs)
        self.clusters.toggle = True
        self.clusters.processing_from = True
        self.get_fluxMatrix(*self._fluxMatrixParams, use_ray=use_ray, progress_bar=progress_bar)
        self.clusters.processing_from = False
        self.clusters.toggle = False

        fluxMatrix = self.fluxMatrixRaw[pcoord_sort_indices, :][:, pcoord_sort_indices]
        self.fluxMatrix = fluxMatrix / np.sum(fluxMatrix)
        self.fluxMatrixRaw = fmatrix_original

        self.indBasis = np.array([self.n_clusters])
        self.indTargets = np.array([self.n_clusters + 1])
        self.nBins = self.n_clusters + 2
        self.update_sorted_cluster_centers()
        self.cluster_mapping = {x: x for x in range(self.n_clusters + 2)}

        fmatrix = self.fluxMatrix.copy()
        fmatrix[-1, -2] = 1.0
        sets = connected_sets(fmatrix, directed=True)
        log.debug(f"After cleaning, shape is {fmatrix.shape} and disconnected sets are: {sets[1:]}")
        assert len(sets[start_cleaning_idx:]) == 0, "Still not clean after cleaning!"

    # ------------------------------------------------------------------------------------------
    def _stacked_dtrajs(self, dtrajs=None):
        dtrajs = self.dtrajs if dtrajs is None else dtrajs
        if len(dtrajs) == 0:
            return np.zeros(0, dtype=np.int64)
        return np.ascontiguousarray(np.concatenate([np.asarray(d, dtype=np.int64) for d in dtrajs]))

    def get_cluster_centers(self):
        """reference: _clustering.py:1528-1599.  Mean / min / max of pcoord dimension 0 over the members of every
        cluster, then the permutation that sorts the clusters by that mean.

        The reference loops over clusters and, for each, runs ``np.where`` over every iteration's dtraj
        (O(n_clusters x frames)); here the stacked dtrajs are grouped by label once on the device and every label is
        reduced by one warp.  Quirks kept: only pcoord dimension 0 is read (``pcoordSet[idx, 0]``) and broadcast over
        the pcoord dimensions; the basis / target rows take ``self.basis_bin_center`` / ``self.target_bin_center``
        (singular: attributes the reference initialises to None and never sets -> NaN rows, which argsort puts
        last); clusters without members get NaN and a warning."""
        import torch

        from .. import ops
        from ..engine import require_cuda

        n = int(self.n_clusters)
        P = self.pcoord_ndim
        centers = np.zeros((n + 2, P))
        ranges = np.zeros((n + 2, P, 2))
        centers[n + 1] = getattr(self, "target_bin_center", None)
        centers[n] = getattr(self, "basis_bin_center", None)
        ranges[n + 1] = [getattr(self, "target_bin_center", None)] * 2
        ranges[n] = [getattr(self, "basis_bin_center", None)] * 2

        labels = self._stacked_dtrajs()
        N = labels.shape[0]
        if self.pcoordSet is None or self.pcoordSet.shape[0] < N:
            raise IndexError(f"pcoordSet holds {0 if self.pcoordSet is None else self.pcoordSet.shape[0]} segments but "
                             f"the dtrajs hold {N}; run get_coordSet() over the discretized iterations first")
        dev = require_cuda()
        lab_d = torch.from_numpy(labels).to(dev)
        val_d = torch.from_numpy(np.ascontiguousarray(self.pcoordSet[:N, 0])).to(dev)
        members, seg_start = ops.group_by_label(lab_d, n)
        count, total, vmin, vmax = ops.label_stats(val_d, members, seg_start, n)
        members_h = members.cpu().numpy()
        seg_h = seg_start.cpu().numpy()
        count_h, total_h, vmin_h, vmax_h = (t.cpu().numpy() for t in (count, total, vmin, vmax))

        group_size = np.diff(seg_h[: n + 1])
        empty = group_size == 0
        for cluster in np.flatnonzero(empty):
            log.warning(f"No trajectories in cluster {cluster}! (Target was {n + 1})")
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = total_h / count_h                  # members that are all NaN: 0/0 = NaN, as nanmean gives
        lo = np.where(count_h > 0, vmin_h, np.nan)
        hi = np.where(count_h > 0, vmax_h, np.nan)
        centers[:n] = mean[:, None]
        ranges[:n, :, 0] = lo[:, None]
        ranges[:n, :, 1] = hi[:, None]

        pcoord_sort_indices = np.argsort(centers[:, 0])
        self.targetRMSD_centers = centers[pcoord_sort_indices]
        self.targetRMSD_minmax = ranges[pcoord_sort_indices]
        # per-cluster member pcoords, as views of ONE gathered array (the reference builds a Python list per cluster)
        gathered = self.pcoordSet[:N, 0][members_h[: seg_h[n]]] if N else np.zeros(0)
        per_cluster = np.empty(n + 2, dtype=object)
        for c in range(n):
            per_cluster[c] = gathered[seg_h[c]:seg_h[c + 1]]
        per_cluster[n] = []
        per_cluster[n + 1] = []
        self.targetRMSD_all = per_cluster[pcoord_sort_indices]
        return pcoord_sort_indices

    def update_sorted_cluster_centers(self):
        """reference: _clustering.py:1601-1611."""
        log.info("Note: Sorting bins, assuming that pcoord 0 is meaningful for sorting")
        bin_centers = self.targetRMSD_centers[:, 0]
        bin_centers[self.indTargets] = self.target_bin_centers
        bin_centers[self.indBasis] = self.basis_bin_centers
        self.all_centers = bin_centers
        self.sorted_centers = np.argsort(bin_centers)

    def update_cluster_structures(self, build_pcoord_cache=False):
        """reference: _clustering.py:1398-1526.  ``cluster_structures[c]`` / ``cluster_structure_weights[c]``: the
        end structures and WE weights of every segment discretized into cluster ``c`` over iterations
        ``1 .. maxIter-2``, in (iteration, segment) order.  The grouping is one device group-by-label of the stacked
        dtrajs; the structures are then gathered per cluster."""
        import torch

        from .. import ops
        from ..engine import require_cuda

        assert self.clusters is not None, "Clusters have not been computed!"
        last = self.maxIter - 1
        iters = list(range(1, last))
        seg_counts = [int(self.numSegments[it - 1]) for it in iters]
        labels = self._stacked_dtrajs([np.asarray(self.dtrajs[it - 1])[:s] for it, s in zip(iters, seg_counts)])
        N = labels.shape[0]
        coords, weights, pcoords, where = [], [], [], []
        for it, s in zip(iters, seg_counts):
            if it not in self.seg_weights.keys():
                self.load_iter_data(it)
            c = self.get_iter_coordinates(it)
            coords.append(c[:s])
            weights.append(np.asarray(self.seg_weights[it])[:s])
            pcoords.append(self.pcoord1List[:s])
            where.append(np.stack([np.full(s, it), self.segindList[:s], self.westList[:s]], axis=1))
        if N and any(int(lbl) in self.removed_clusters for lbl in np.unique(labels)):
            raise Exception("This dtraj point was in a removed cluster -- this should never happen!")
        n_labels = int(labels.max()) + 1 if N else 1
        dev = require_cuda()
        members, seg_start = ops.group_by_label(torch.from_numpy(labels).to(dev), n_labels)
        members_h, seg_h = members.cpu().numpy(), seg_start.cpu().numpy()
        all_coords = np.concatenate(coords) if coords else np.zeros((0, 0, 0))
        all_weights = np.concatenate(weights) if weights else np.zeros(0)
        all_pcoords = np.concatenate(pcoords) if pcoords else np.zeros((0, self.pcoord_ndim))
        all_where = np.concatenate(where) if where else np.zeros((0, 3), dtype=np.int64)
        cluster_structures, cluster_structure_weights, structure_iteration_segments = {}, {}, {}
        pcoord_cache = {} if build_pcoord_cache else None
        # dictionary keys in order of first appearance, as the reference's incremental construction gives
        _, first_pos = np.unique(labels, return_index=True)
        for c in labels[np.sort(first_pos)]:
            idx = members_h[seg_h[c]:seg_h[c + 1]]
            c = int(c)
            cluster_structures[c] = list(all_coords[idx])
            cluster_structure_weights[c] = list(all_weights[idx])
            structure_iteration_segments[c] = [[int(a), int(b), self.fileList[int(f)] if self.fileList else int(f)]
                                               for a, b, f in all_where[idx]]
            if build_pcoord_cache:
                pcoord_cache[c] = list(all_pcoords[idx])
        assert len(cluster_structures) == len(cluster_structure_weights), "Structures and weights have different numbers of bins?"
        self.cluster_structures = cluster_structures
        self.cluster_structure_weights = cluster_structure_weights
        self.pcoord_cache = pcoord_cache
        self.structure_iteration_segments = structure_iteration_segments
        log.debug("Cluster structure mapping completed.")
