"""``StratifiedClusters`` with the reference's constructor, attributes and ``predict`` semantics, running
on the GPU.

reference: msm_we/stratified_clustering.py:6-212.  One independent k-means model per WE bin;
``predict`` maps every coordinate to its (remapped) WE bin, then to the nearest centre of that bin, and
adds the bin's offset; points whose pcoord lies in the target / basis get ``T+1`` / ``T``
(``T`` = number of centres over all fitted bins).

``BinClusterModel`` stands in for ``sklearn.cluster.MiniBatchKMeans`` for the subset of its interface
msm_we touches (``cluster_centers_`` get/set/``hasattr``, ``partial_fit``, ``predict`` and the fitted
state attributes).  Its arithmetic is the K1/K2 kernels; its host logic (initialisation, random
reassignment, counters) restates sklearn/cluster/_kmeans.py:1932-1957, 2039-2054, 2227-2325 and
:1651-1682 so that, given the same ``random_state``, it draws the same random numbers in the same
order as sklearn does.
"""
from __future__ import annotations

import numbers
import zlib

import numpy as np

from ._logging import log


def check_random_state(seed):
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, numbers.Integral):
        return np.random.RandomState(seed)
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError(f"{seed!r} cannot be used to seed a numpy.random.RandomState instance")


def _sq_euclidean_to(X, x_sq, Y):
    """``||y - x||^2`` for every row y of Y and x of X, the way sklearn's ``_euclidean_distances`` forms it
    (-2 Y.X^T, += ||y||^2, += ||x||^2, clip at 0) so k-means++ picks the same candidates."""
    d = -2.0 * (Y @ X.T)
    d += np.einsum("ij,ij->i", Y, Y)[:, None]
    d += x_sq[None, :]
    np.maximum(d, 0, out=d)
    return d


def kmeans_plusplus(X, n_clusters, x_sq, sample_weight, random_state, n_local_trials=None):
    """k-means++ seeding, sklearn/cluster/_kmeans.py `_kmeans_plusplus` (host; runs once per WE bin on at
    most ``init_size`` rows)."""
    n_samples, n_features = X.shape
    centers = np.empty((n_clusters, n_features), dtype=X.dtype)
    if n_local_trials is None:
        n_local_trials = 2 + int(np.log(n_clusters))
    center_id = random_state.choice(n_samples, p=sample_weight / sample_weight.sum())
    centers[0] = X[center_id]
    closest = _sq_euclidean_to(X, x_sq, centers[0, np.newaxis])
    current_pot = closest @ sample_weight
    for c in range(1, n_clusters):
        rand_vals = random_state.uniform(size=n_local_trials) * current_pot
        candidate_ids = np.searchsorted(np.cumsum(sample_weight * closest), rand_vals)
        np.clip(candidate_ids, None, closest.size - 1, out=candidate_ids)
        dist = _sq_euclidean_to(X, x_sq, X[candidate_ids])
        np.minimum(closest, dist, out=dist)
        pots = dist @ sample_weight.reshape(-1, 1)
        best = int(np.argmin(pots))
        current_pot = pots[best]
        closest = dist[best][np.newaxis, :]
        centers[c] = X[candidate_ids[best]]
    return centers


class BinClusterModel:
    """MiniBatchKMeans-shaped per-bin model (see module docstring)."""

    def __init__(self, n_clusters=8, *, init="k-means++", max_iter=100, batch_size=1024, verbose=0,
                 compute_labels=True, random_state=None, tol=0.0, max_no_improvement=10, init_size=None,
                 n_init="auto", reassignment_ratio=0.01, **extra):
        self.n_clusters = n_clusters
        self.init = init
        self.max_iter = max_iter
        self.batch_size = batch_size
        self.verbose = verbose
        self.compute_labels = compute_labels
        self.random_state = random_state
        self.tol = tol
        self.max_no_improvement = max_no_improvement
        self.init_size = init_size
        self.n_init = n_init
        self.reassignment_ratio = reassignment_ratio
        # GPU knobs ride in **_cluster_args: accepted and ignored here
        self.extra = dict(extra)

    # ---- sklearn-compatible bookkeeping --------------------------------------------------------
    def _check_params_vs_input(self, X):
        if X.shape[0] < self.n_clusters:
            raise ValueError(f"n_samples={X.shape[0]} should be >= n_clusters={self.n_clusters}.")
        self._batch_size = min(self.batch_size, X.shape[0])
        self._init_size = self.init_size
        if self._init_size is None:
            self._init_size = 3 * self._batch_size
            if self._init_size < self.n_clusters:
                self._init_size = 3 * self.n_clusters
        elif self._init_size < self.n_clusters:
            self._init_size = 3 * self.n_clusters
        self._init_size = min(self._init_size, X.shape[0])
        if self.reassignment_ratio < 0:
            raise ValueError(f"reassignment_ratio should be >= 0, got {self.reassignment_ratio} instead.")

    def _init_centroids(self, X, x_sq, sample_weight):
        rs = self._random_state
        n_samples = X.shape[0]
        init = self.init
        if self._init_size is not None and self._init_size < n_samples:
            idx = rs.randint(0, n_samples, self._init_size)
            X, x_sq, sample_weight = X[idx], x_sq[idx], sample_weight[idx]
            n_samples = X.shape[0]
        if isinstance(init, str) and init == "k-means++":
            return kmeans_plusplus(X, self.n_clusters, x_sq, sample_weight, rs)
        if isinstance(init, str) and init == "random":
            seeds = rs.choice(n_samples, size=self.n_clusters, replace=False, p=sample_weight / sample_weight.sum())
            return X[seeds].copy()
        if callable(init):
            return np.ascontiguousarray(init(X, self.n_clusters, random_state=rs), dtype=np.float64)
        c = np.array(init, dtype=np.float64, order="C", copy=True)
        if c.shape != (self.n_clusters, X.shape[1]):
            raise ValueError(f"The shape of the initial centers {c.shape} does not match "
                             f"(n_clusters, n_features) = {(self.n_clusters, X.shape[1])}.")
        return c

    def _random_reassign(self):
        self._n_since_last_reassign += self._batch_size
        if (self._counts == 0).any() or self._n_since_last_reassign >= (10 * self.n_clusters):
            self._n_since_last_reassign = 0
            return True
        return False

    def _prepare(self, X, sample_weight):
        """Host half of partial_fit before the step; returns (X, w, random_reassign)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("Expected 2D array")
        has_centers = hasattr(self, "cluster_centers_")
        if has_centers and X.shape[1] != self.cluster_centers_.shape[1]:
            raise ValueError(f"X has {X.shape[1]} features, but the model is expecting "
                             f"{self.cluster_centers_.shape[1]} features as input.")
        self._random_state = getattr(self, "_random_state", None) or check_random_state(self.random_state)
        if sample_weight is None:
            w = np.ones(X.shape[0], dtype=np.float64)
        else:
            w = np.ascontiguousarray(sample_weight, dtype=np.float64)
            if w.shape != (X.shape[0],):
                raise ValueError("sample_weight.shape == {}, expected {}!".format(w.shape, (X.shape[0],)))
        self.n_steps_ = getattr(self, "n_steps_", 0)
        if not has_centers:
            self._check_params_vs_input(X)
            self.n_features_in_ = X.shape[1]
            x_sq = np.einsum("ij,ij->i", X, X)
            self.cluster_centers_ = self._init_centroids(X, x_sq, w)
            self._counts = np.zeros(self.n_clusters, dtype=np.float64)
            self._n_since_last_reassign = 0
        return X, w, self._random_reassign()

    def _finish(self, X, random_reassign):
        """Host half after the GPU step: random reassignment of low-count centres
        (sklearn/cluster/_kmeans.py:1651-1682) and counters."""
        if random_reassign and self.reassignment_ratio > 0:
            counts = self._counts
            to_reassign = counts < self.reassignment_ratio * counts.max()
            if to_reassign.sum() > 0.5 * X.shape[0]:
                keep = np.argsort(counts)[int(0.5 * X.shape[0]):]
                to_reassign[keep] = False
            n_reassigns = to_reassign.sum()
            if n_reassigns:
                new_centers = self._random_state.choice(X.shape[0], replace=False, size=n_reassigns)
                self.cluster_centers_[to_reassign] = X[new_centers]
            counts[to_reassign] = np.min(counts[~to_reassign])
        self.n_steps_ += 1
        self._n_features_out = self.cluster_centers_.shape[0]

    # ---- public ---------------------------------------------------------------------------------
    def partial_fit(self, X, y=None, sample_weight=None):
        from .clustering_ops import partial_fit_models

        partial_fit_models([(self, X, sample_weight)])
        return self

    def predict(self, X):
        from .clustering_ops import predict_single

        if not hasattr(self, "cluster_centers_"):
            raise AttributeError("This BinClusterModel instance is not fitted yet.")
        return predict_single(self.cluster_centers_, X)

    def __repr__(self):
        return f"BinClusterModel(n_clusters={self.n_clusters}, fitted={hasattr(self, 'cluster_centers_')})"


_HASH_POOL = None


def _hash_pool():
    global _HASH_POOL
    if _HASH_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor

        n = min(8, max(1, (os.cpu_count() or 2) // 2))
        _HASH_POOL = ThreadPoolExecutor(max_workers=n, thread_name_prefix="mwe-hash") if n > 1 else False
    return _HASH_POOL or None


class StratifiedClusters:
    """reference: msm_we/stratified_clustering.py:6-212 (same constructor, attributes and methods)."""

    def __init__(self, bin_mapper, model, n_clusters, target_bins, **_cluster_args):
        cluster_args = {"n_clusters": n_clusters, "max_iter": 100}
        cluster_args.update(_cluster_args)
        self.n_clusters_per_bin = n_clusters
        self.bin_mapper = bin_mapper
        self.n_total_clusters = self.n_clusters_per_bin * (self.bin_mapper.nbins - len(target_bins))
        log.info(f"Doing stratified clustering with {self.n_total_clusters} total clusters")
        self.cluster_args = cluster_args
        self.model = model
        self.cluster_models = [BinClusterModel(**cluster_args) for _ in range(self.bin_mapper.nbins)]
        self.processing_from = False
        self.toggle = False
        self.we_remap = {x: x for x in range(self.bin_mapper.nbins)}
        self.legitimate_bins = range(self.bin_mapper.nbins)
        self.target_bins = set()
        self.basis_bins = set()
        self._device = None
        self._device_key = None

    # GPU handles are caches: drop them when pickling / deep-copying
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_device"] = None
        state["_device_key"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.__dict__.setdefault("_device", None)
        self.__dict__.setdefault("_device_key", None)

    def centers_per_bin(self):
        return [getattr(m, "cluster_centers_", None) for m in self.cluster_models]

    def _fingerprint(self):
        # the centres are small (K x D per bin): hash ALL their bytes, so an in-place edit is never missed.  20 MB at
        # config-5 size: zlib releases the GIL, so the per-bin hashes run on a few threads (6 ms -> 1 ms)
        def one(m):
            c = getattr(m, "cluster_centers_", None)
            if c is None:
                return None
            c = np.asarray(c)
            return (c.shape, zlib.crc32(np.ascontiguousarray(c).view(np.uint8).reshape(-1)))

        models = self.cluster_models
        nbytes = sum(getattr(getattr(m, "cluster_centers_", None), "nbytes", 0) for m in models)
        pool = _hash_pool() if nbytes > (1 << 20) else None
        parts = list(pool.map(one, models)) if pool is not None else [one(m) for m in models]
        model = self.model
        bounds = (np.asarray(model.basis_pcoord_bounds).tobytes(), np.asarray(model.target_pcoord_bounds).tobytes())
        return (tuple(parts), tuple(sorted(self.we_remap.items())), id(self.bin_mapper), bounds)

    def device_state(self):
        """Device snapshot, rebuilt whenever centres, ``we_remap``, the mapper or the bounds changed."""
        from .engine import DeviceClusters

        key = self._fingerprint()
        if self._device is None or key != self._device_key:
            self._device = DeviceClusters(self.bin_mapper, self.centers_per_bin(), self.we_remap,
                                          self.model.basis_pcoord_bounds, self.model.target_pcoord_bounds,
                                          self.model.pcoord_ndim)
            self._device_key = key
        return self._device

    def adopt_device_centers(self, centers_dev):
        """The caller refined the centres ON the device and has already written the same values to every
        ``cluster_centers_``: keep the device snapshot (same shapes) instead of re-uploading them on the next use."""
        from . import ops

        dev = self._device
        if dev is None or tuple(centers_dev.shape) != tuple(dev.centers.shape):
            return
        dev.centers.copy_(centers_dev)
        dev.csq = ops.centers_sqnorm(dev.centers)
        self._device_key = self._fingerprint()

    def predict(self, coords):
        """Same contract as the reference's ``predict`` (stratified_clustering.py:101-212): bins come from
        ``model.pcoord0List`` when ``processing_from`` else ``model.pcoord1List``; returns int64 labels;
        records ``target_bins`` / ``basis_bins``; flips ``processing_from`` when ``toggle`` is set."""
        import torch

        iter_pcoords = self.model.pcoord0List if self.processing_from else self.model.pcoord1List
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        pc = np.ascontiguousarray(iter_pcoords, dtype=np.float64)
        if pc.ndim == 1:
            pc = pc[:, None]
        if coords.ndim != 2 or coords.shape[0] != pc.shape[0]:
            # the reference indexes is_target[i] for every coord and fails with IndexError on a mismatch
            raise IndexError(f"{coords.shape[0]} coordinates but {pc.shape[0]} progress coordinates")
        if coords.shape[0] == 0:
            return np.array([])
        dev = self.device_state()
        X = torch.from_numpy(coords).to(dev.device)
        P = torch.from_numpy(pc).to(dev.device)
        labels, bins, flags = dev.predict(X, P, pcoord_host=pc)
        labels_h = labels.cpu().numpy()
        bins_h = bins.cpu().numpy()
        flags_h = flags.cpu().numpy()
        dev.check_errors()
        is_target = (flags_h & 2) != 0
        is_basis = ((flags_h & 1) != 0) & ~is_target
        self.target_bins.update(int(b) for b in np.unique(bins_h[is_target]))
        self.basis_bins.update(int(b) for b in np.unique(bins_h[is_basis]))
        if self.toggle:
            self.processing_from = not self.processing_from
            log.debug(f"Finished and toggling... Next iteration will use pcoord{not self.processing_from:d}List")
        return labels_h
