"""Clustering + discretization mixin: the reference's method names and control flow, GPU arithmetic.

reference: msm_we/_hamsm/_clustering.py -- ``cluster_coordinates`` (:142-195), ``cluster_stratified``
(:525-746), ``do_stratified_clustering`` (:748-918), ``launch_ray_discretization`` (:1144-1242),
``do_stratified_ray_discretization`` (:1244-1329), ``find_nearest_bin`` (:1331-1396).

What changed underneath:
* no Ray, no fork-per-iteration ``ProcessPoolExecutor``: ``use_ray`` is accepted and ignored;
* ``do_stratified_clustering`` keeps the reference's batching rule (pull iterations until every seen WE
  bin holds >= n_clusters segments) but runs all WE bins of a batch through one K1 + one K2 launch;
* ``launch_ray_discretization`` stages many iterations into one pinned buffer, and K0 + K1 label every
  parent and child frame of the chunk in a single launch sequence instead of one sklearn call per
  segment.
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
    kmeans_model = deepcopy(kmeans_model)
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
    pre_discretization_model = None
    post_cluster_model = None

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
            bin_mapper = getattr(self, "bin_mapper", None)
            if bin_mapper is None:
                raise Exception("No bin mapper: the reference unpickles it from the WESTPA HDF5 file with westpa, "
                                "which is not available here; pass user_bin_mapper=")
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

    # ------------------------------------------------------------------------------------------
    def do_stratified_clustering(self, arg):
        """reference: _clustering.py:748-918.  Returns ``(kmeans_models, used_iters, unique_bins,
        unfilled_bins)``.  The per-bin ``partial_fit`` calls of the reference (:890-916) are issued as ONE
        batched GPU step over all bins of the batch."""
        from ..clustering_ops import partial_fit_models

        self, kmeans_models, iters_to_use, processCoordinates, ignored_bins = arg
        iters_to_use = list(iters_to_use)
        bin_mapper = kmeans_models.bin_mapper
        min_coords = kmeans_models.cluster_args["n_clusters"]
        all_bins_have_segments = False
        used_iters = -1
        iter_coords = []
        seg_weights = None
        pcoords = []
        unique_bins = np.array([])
        counts = np.array([])
        we_bin_assignments = np.array([])
        unfilled_bins = []
        iteration = None

        while not all_bins_have_segments:
            unfilled_bins = []
            try:
                iteration = iters_to_use.pop(0)
            except IndexError:
                log.warning(f"At iteration {iteration} (pulled {used_iters} extra), couldn't get segments in all "
                            f"bins, and no iterations left.")
                unfilled_bins = unique_bins[counts < min_coords]
                filled_bins = np.setdiff1d(unique_bins, unfilled_bins)
                for unfilled_bin in unfilled_bins:
                    nearest_filled_bin = self.find_nearest_bin(bin_mapper, unfilled_bin, list(filled_bins))
                    unfilled_bin_indices = np.where(we_bin_assignments == unfilled_bin)
                    log.warning(f"Remapping {len(unfilled_bin_indices)} segments from unfilled bin {unfilled_bin} to "
                                f"{nearest_filled_bin} for stratified clustering")
                    we_bin_assignments[unfilled_bin_indices] = nearest_filled_bin
                unique_bins = filled_bins
                break

            if iteration > self.maxIter:
                log.warning(f"At iteration {iteration} (pulled {used_iters} extra), couldn't get segments in all "
                            f"bins, and no iterations left")
                break

            used_iters += 1
            _iter_coords = self.get_iter_coordinates(iteration)
            _seg_weights = self.seg_weights[iteration]
            if used_iters == 0:
                iter_coords = _iter_coords
                seg_weights = _seg_weights
                pcoords = [x for x in self.pcoord0List]
            else:
                iter_coords = np.append(iter_coords, _iter_coords, axis=0)
                pcoords.extend(self.pcoord0List)
                if self.use_weights_in_clustering:
                    seg_weights = np.append(seg_weights, _seg_weights, axis=0)

            pcoord_array = np.array(pcoords)
            assert pcoord_array.shape[0] == iter_coords.shape[0], f"{pcoord_array.shape}, {iter_coords.shape}"

            # segments whose PARENT pcoord is in the basis or target are ignored (:872-874)
            pcoord_is_target = self.is_WE_target(pcoord_array)
            pcoord_is_basis = self.is_WE_basis(pcoord_array)
            pcoord_array = pcoord_array[~(pcoord_is_target | pcoord_is_basis)]
            if len(pcoord_array) > 0:
                we_bin_assignments = np.asarray(bin_mapper.assign(pcoord_array))
            else:
                we_bin_assignments = np.array([])
            unique_bins, counts = np.unique(we_bin_assignments, return_counts=True)
            all_bins_have_segments = np.all(counts >= min_coords)

        # one batched GPU step over every bin that received segments.  NOTE: as in the reference
        # (:892-899) the row indices computed on the basis/target-FILTERED pcoord array index the
        # UNFILTERED coordinates and weights; reproduced deliberately.
        batch = []
        for _bin in unique_bins:
            segs_in_bin = np.argwhere(we_bin_assignments == _bin)
            transformed_coords = self.coordinates.transform(processCoordinates(np.squeeze(iter_coords[segs_in_bin])))
            weights = seg_weights[segs_in_bin].squeeze() if self.use_weights_in_clustering else None
            batch.append((kmeans_models.cluster_models[int(_bin)], transformed_coords, weights))
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

        self.check_connect_ray()
        self.dtrajs = []
        if self.pre_discretization_model is None:
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
        row_bytes = 2 * (D + P) * 8
        seg_counts = [int(self.numSegments[it - 1]) if self.numSegments is not None and it <= len(self.numSegments)
                      else None for it in range(1, self.maxIter)]

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

        stream = torch.cuda.current_stream()
        slots = [dict(event=None, host=None, n=0), dict(event=None, host=None, n=0)]
        inflight = []   # (chunk, n, labels_host, bins_host, flags_host, done_event)

        def stage(ci, chunk):
            from .._pinning import PINS

            n = sum(s for _, s in chunk)
            slot = slots[ci % 2]
            if slot["event"] is not None:
                slot["event"].synchronize()          # the H2D that last read this buffer has finished
            featurise, transform = self.processCoordinates, self.coordinates.transform
            # a fitted linear projection (the reference's PCA ``coordinates.transform``) is applied on the device:
            # the featurised frames are shipped as they are
            projection = getattr(self.coordinates, "device_projection", None)
            projection = projection(dev.device) if callable(projection) else None
            pool = _staging_pool()

            def featurise_pair(item):
                parent_coords, child_coords = self.iter_coordinate_pair(item[0])
                fp, fc = featurise(parent_coords), featurise(child_coords)
                if projection is None:
                    fp, fc = transform(fp), transform(fc)
                return np.asarray(fp), np.asarray(fc)

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

            def fill(dst, feat):
                np.copyto(dst, feat)

            pos, jobs, staged = 0, [], []
            for (it, s), (fp, fc) in zip(chunk, feats):
                rec = self._record(it)
                hp[pos:pos + s] = rec.pcoord0[:, :P]
                hp[n + pos:n + pos + s] = rec.pcoord1[:, :P]
                for off, feat in ((pos, fp), (n + pos, fc)):
                    if feat.shape != (s, Din):
                        raise ValueError(f"featurised coordinates of iteration {it} have shape {feat.shape}, expected {(s, Din)}")
                    if PINS.ensure(feat):
                        # the model's own array is page-locked: straight to the device at link speed
                        X[off:off + s].copy_(torch.from_numpy(feat), non_blocking=True)
                    else:
                        # featurised / projected on the fly (or not lockable): stage through the pinned rows
                        if pool is None:
                            fill(hx[off:off + s], feat)
                        else:
                            jobs.append(pool.submit(fill, hx[off:off + s], feat))
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
            if projection is not None:
                from .. import ops as _ops

                X = _ops.project(X, projection[0], projection[1])
            Pc = hp_t.to(dev.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            slot["event"] = ev
            labels, bins, flags = dev.predict(X, Pc, pcoord_host=hp)
            lh = torch.empty(2 * n, dtype=torch.int64, pin_memory=True)
            bh = torch.empty(2 * n, dtype=torch.int32, pin_memory=True)
            fh = torch.empty(2 * n, dtype=torch.uint8, pin_memory=True)
            lh.copy_(labels, non_blocking=True); bh.copy_(bins, non_blocking=True); fh.copy_(flags, non_blocking=True)
            done = torch.cuda.Event()
            done.record(stream)
            inflight.append((chunk, n, lh, bh, fh, done))

        with ProgressBar(progress_bar) as progress:
            task = progress.add_task(description="Discretizing trajectories", total=n_iters)
            def collect(item):
                chunk, n, lh, bh, fh, done = item
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
                pos = 0
                for it, s in chunk:
                    dtrajs[it - 1] = child_all[pos:pos + s]
                    pair_dtrajs[it - 1] = pairs_all[pos:pos + s]
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

        self.dtrajs = [d for d in dtrajs if d is not None]
        self.pair_dtrajs = [d for d in pair_dtrajs if d is not None]
        log.debug("Discretization complete")

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
