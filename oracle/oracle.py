"""CPU oracle for the msm_we discretization + flux hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``msm_we_b200`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and only as the checker (or as the timed CPU arm), never as the product.

What it is: a plain numpy / scipy / scikit-learn restatement, function by function, of the reference
algorithm (jdrusso/msm_we).  The reference itself cannot be imported in this image (mdtraj, ray,
westpa, h5py, deeptime are absent), and the arithmetic of the path lives in third-party packages:

* scikit-learn (reference pin ``>=0.24,<1.1``, fixtures built with 1.0.2; installed here: 1.9.0) --
  ``MiniBatchKMeans.predict / partial_fit`` and ``KMeans`` Lloyd iterations.  The oracle calls the
  *installed* sklearn where the reference calls sklearn, and also restates the published Cython
  kernels (``sklearn/cluster/_k_means_lloyd.pyx:168-218``, ``_k_means_minibatch.pyx:59-111``,
  ``_kmeans.py:1566-1684, 2039-2054, 2227-2325``) in numpy so both can be compared.
* scipy.sparse (``coo_matrix`` duplicate-summing scatter) -- called directly, as the reference does.
* westpa ``RectilinearBinMapper.assign`` / ``VoronoiBinMapper.assign`` -- NOT installed; restated
  from its published semantics (float32 coordinates, ``lower <= x < upper``, row-major bin index,
  ValueError when out of range).  UNVERIFIED against westpa source.

Parity pinning (see tests/test_oracle_golden.py, tests/golden/make_golden.py):
  * colour-augmented scatter vs the reference's own known-answer test
    (/root/reference/tests/test_non_markov_model.py:8-26);
  * flux scatter sparsity + label conventions vs the reference's pickled ``clustered.obj`` /
    ``fluxmatrix_raw.npy`` (values cannot be pinned: west.h5 with the weights is stripped);
  * assignment / minibatch / Lloyd arithmetic vs the installed scikit-learn, run live.
Label parity and flux *values* are therefore "pinned against sklearn/scipy run here", not against
reference-held golden outputs; DESIGN.md says the same.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix

# --------------------------------------------------------------------------------------------
# a8: basis / target tests            reference: msm_we/msm_we.py:462-527
# --------------------------------------------------------------------------------------------


def is_we_region(pcoords: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    """Strict ``lo < p < hi`` on every pcoord dimension, AND over dimensions.

    reference: modelWE.is_WE_basis / is_WE_target, msm_we/msm_we.py:462-527 (identical logic for the
    two; only the bounds array differs).
    """
    pcoords = np.asarray(pcoords, dtype=np.float64)
    bounds = np.asarray(bounds, dtype=np.float64)
    if pcoords.ndim == 1:
        pcoords = pcoords[:, None]
    inside = np.ones(pcoords.shape[0], dtype=bool)
    for d in range(bounds.shape[0]):
        inside &= (pcoords[:, d] > bounds[d, 0]) & (pcoords[:, d] < bounds[d, 1])
    return inside


# --------------------------------------------------------------------------------------------
# WE bin lookup    reference call sites: stratified_clustering.py:134, _clustering.py:877
# (westpa.core.binning.RectilinearBinMapper / VoronoiBinMapper; restated, UNVERIFIED)
# --------------------------------------------------------------------------------------------


class RectilinearBinMapperOracle:
    """westpa RectilinearBinMapper semantics: float32 coords and boundaries, per dimension
    ``b[i] <= x < b[i+1]``, row-major index with the last dimension fastest, ValueError when a
    coordinate falls outside the bin space."""

    def __init__(self, boundaries):
        self.boundaries = [np.asarray(b, dtype=np.float32) for b in boundaries]
        self.ndim = len(self.boundaries)
        self.nbins = int(np.prod([len(b) - 1 for b in self.boundaries]))

    def assign(self, coords):
        coords = np.asarray(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords[:, None]
        c32 = coords.astype(np.float32)
        index = np.zeros(coords.shape[0], dtype=np.int64)
        for d, b in enumerate(self.boundaries):
            # number of boundaries <= x, minus one  ==  i such that b[i] <= x < b[i+1]
            pos = np.searchsorted(b, c32[:, d], side="right") - 1
            bad = (pos < 0) | (pos >= len(b) - 1) | np.isnan(c32[:, d])
            if bad.any():
                raise ValueError("coordinate outside of bin space")
            index = index * (len(b) - 1) + pos
        return index


class VoronoiBinMapperOracle:
    """westpa VoronoiBinMapper with the Euclidean ``dfunc``: nearest centre, first minimum wins."""

    def __init__(self, centers):
        self.centers = np.asarray(centers, dtype=np.float32)
        if self.centers.ndim == 1:
            self.centers = self.centers[:, None]
        self.nbins = self.centers.shape[0]
        self.ndim = self.centers.shape[1]

    def assign(self, coords):
        coords = np.asarray(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords[:, None]
        c32 = coords.astype(np.float32)
        # float32 squared distances accumulated in dimension order (the GPU helper does the same)
        d2 = np.zeros((c32.shape[0], self.nbins), dtype=np.float32)
        for d in range(self.ndim):
            diff = c32[:, d : d + 1] - self.centers[None, :, d]
            d2 = d2 + diff * diff
        return np.argmin(d2, axis=1).astype(np.int64)


# --------------------------------------------------------------------------------------------
# K-means E step    reference arithmetic: sklearn/cluster/_k_means_lloyd.pyx:168-218
# --------------------------------------------------------------------------------------------


def kmeans_scores(X: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """``||c||^2 - 2 x.c`` (sklearn drops ``||x||^2``), fp64.  _k_means_lloyd.pyx:193-203."""
    X = np.asarray(X, dtype=np.float64)
    centers = np.asarray(centers, dtype=np.float64)
    csq = np.einsum("ij,ij->i", centers, centers)
    return csq[None, :] - 2.0 * (X @ centers.T)


def kmeans_assign(X: np.ndarray, centers: np.ndarray, return_margin: bool = False):
    """Nearest centre with strict ``<`` (lowest index wins ties).  _k_means_lloyd.pyx:205-213.

    ``return_margin`` also returns, per point, (second-best - best) / scale, with scale the
    magnitude of the terms entering the subtraction: a point with a margin below ~1e-12 is a
    floating-point near-tie whose label legitimately depends on BLAS summation order.
    """
    s = kmeans_scores(X, centers)
    labels = np.argmin(s, axis=1)  # first occurrence == strict-< scan from index 0
    if not return_margin:
        return labels
    if s.shape[1] == 1:
        return labels, np.full(s.shape[0], np.inf)
    part = np.partition(s, 1, axis=1)
    gap = part[:, 1] - part[:, 0]
    xn = np.sqrt(np.einsum("ij,ij->i", X, X))
    cn = np.sqrt(np.einsum("ij,ij->i", centers, centers)).max()
    scale = 2.0 * xn * cn + cn * cn + 1e-300
    return labels, gap / scale


TIE_C = 4.0


def tie_tolerance(X, centers):
    """Per-point width of the band inside which two scores count as tied on the GPU:
    TIE_C (D+8) 2^-53 cmax (2 ||x|| + cmax), cmax = max_j ||c_j||  -- a few times the worst-case rounding
    error of evaluating ``||c||^2 - 2 x.c`` in fp64 (msm_we_b200/csrc/assign.cu, assign_recheck_kernel)."""
    X = np.asarray(X, dtype=np.float64)
    centers = np.asarray(centers, dtype=np.float64)
    D = X.shape[1]
    xn = np.sqrt(np.einsum("ij,ij->i", X, X))
    cmax = np.sqrt(np.einsum("ij,ij->i", centers, centers).max())
    return TIE_C * (D + 8) * 2.0 ** -53 * cmax * (2.0 * xn + cmax)


def kmeans_assign_tiebreak(X, centers, return_ambiguous=False):
    """The GPU path's DEFINED resolution of floating-point near-ties: the lowest index among the centres
    whose score is within ``tie_tolerance`` of the minimum.  Outside that band it is identical to
    ``kmeans_assign`` (the reference's strict-< scan); inside it the reference's own answer depends on
    the BLAS summation order, except for exactly equal scores where both pick the lowest index.
    ``return_ambiguous`` marks rows where some score lies so close to the band's edge (0.5..1.5 tol)
    that evaluation order could move it across."""
    s = kmeans_scores(X, centers)
    tol = tie_tolerance(X, centers)
    smin = np.min(s, axis=1)
    with np.errstate(invalid="ignore"):
        rel = s - smin[:, None]
        tied = rel <= tol[:, None]
    labels = np.argmax(tied, axis=1)  # first True; an all-NaN row gives 0 like the reference's scan
    if not return_ambiguous:
        return labels
    with np.errstate(invalid="ignore"):
        amb = ((rel > 0.5 * tol[:, None]) & (rel < 1.5 * tol[:, None])).any(axis=1)
    return labels, amb


def kmeans_assign_exact(X, centers):
    """Slow exact-rational restatement for tiny tie-break cases (pure Python, fractions)."""
    from fractions import Fraction

    X = np.asarray(X, dtype=np.float64)
    centers = np.asarray(centers, dtype=np.float64)
    out = np.zeros(X.shape[0], dtype=np.int64)
    for i in range(X.shape[0]):
        best, bestj = None, 0
        for j in range(centers.shape[0]):
            s = Fraction(0)
            for k in range(X.shape[1]):
                c = Fraction(float(centers[j, k]))
                s += c * c - 2 * Fraction(float(X[i, k])) * c
            if best is None or s < best:
                best, bestj = s, j
        out[i] = bestj
    return out


def make_fitted_minibatch(centers, **cluster_args):
    """A real sklearn ``MiniBatchKMeans`` carrying given centres, so ``.predict`` works without a fit
    (what unpickling a fitted model gives the reference)."""
    from sklearn.cluster import MiniBatchKMeans

    centers = np.ascontiguousarray(centers, dtype=np.float64)
    args = {"n_clusters": centers.shape[0], "max_iter": 100}
    args.update(cluster_args)
    m = MiniBatchKMeans(**args)
    m.cluster_centers_ = centers
    m._n_threads = 1
    m.n_features_in_ = centers.shape[1]
    m._n_features_out = centers.shape[0]
    m._counts = np.zeros(centers.shape[0])
    m._n_since_last_reassign = 0
    return m


# --------------------------------------------------------------------------------------------
# a1/a2: StratifiedClusters.predict      reference: msm_we/stratified_clustering.py:101-212
# --------------------------------------------------------------------------------------------


class StratifiedOracle:
    """Holds exactly the state ``StratifiedClusters`` holds (stratified_clustering.py:21-99)."""

    def __init__(self, bin_mapper, centers_per_bin, basis_bounds, target_bounds, we_remap=None):
        self.bin_mapper = bin_mapper
        # list over WE bins; None == "no cluster_centers_ attribute" (never fitted)
        self.centers_per_bin = [None if c is None else np.asarray(c, dtype=np.float64) for c in centers_per_bin]
        self.basis_bounds = np.asarray(basis_bounds, dtype=np.float64)
        self.target_bounds = np.asarray(target_bounds, dtype=np.float64)
        nb = len(self.centers_per_bin)
        self.we_remap = {b: b for b in range(nb)} if we_remap is None else dict(we_remap)
        self.target_bins = set()
        self.basis_bins = set()

    # offsets: stratified_clustering.py:143-150 and :178-185
    def _sizes(self):
        return [0 if c is None else len(c) for c in self.centers_per_bin]

    def predict(self, coords, pcoords, literal=False, models=None):
        """Labels for ``coords`` binned by ``pcoords`` (the caller picks pcoord0List or pcoord1List,
        stratified_clustering.py:129-132).

        ``literal=True`` follows the reference loop to the letter: one sklearn
        ``predict([coord])`` per segment (stratified_clustering.py:152-203).  Otherwise one numpy
        E-step per WE bin, which gives the same labels.
        """
        coords = np.asarray(coords, dtype=np.float64)
        pcoords = np.asarray(pcoords, dtype=np.float64)
        if pcoords.ndim == 1:
            pcoords = pcoords[:, None]
        we_bins = np.array([self.we_remap[int(b)] for b in self.bin_mapper.assign(pcoords)], dtype=np.int64)
        is_target = is_we_region(pcoords, self.target_bounds)
        is_basis = is_we_region(pcoords, self.basis_bounds)
        sizes = self._sizes()
        total = int(sum(sizes))
        offsets = np.concatenate([[0], np.cumsum(sizes)])[:-1]
        out = np.zeros(coords.shape[0], dtype=np.int64)

        # target test first, then basis   (stratified_clustering.py:159-169)
        out[is_target] = total + 1
        out[is_basis & ~is_target] = total
        for b in np.unique(we_bins[is_target]):
            self.target_bins.add(int(b))
        for b in np.unique(we_bins[is_basis & ~is_target]):
            self.basis_bins.add(int(b))

        free = ~(is_target | is_basis)
        for b in np.unique(we_bins[free]):
            b = int(b)
            sel = np.where(free & (we_bins == b))[0]
            # stratified_clustering.py:187-189
            assert self.centers_per_bin[b] is not None, f"Not initialized in bin {b}"
            if literal:
                model = models[b] if models is not None else make_fitted_minibatch(self.centers_per_bin[b])
                for i in sel:
                    out[i] = model.predict([coords[i]])[0] + offsets[b]
            else:
                out[sel] = kmeans_assign_tiebreak(coords[sel], self.centers_per_bin[b]) + offsets[b]
        return out


# --------------------------------------------------------------------------------------------
# a5': minibatch centre update     sklearn/cluster/_k_means_minibatch.pyx:59-111 and
#      random reassignment           sklearn/cluster/_kmeans.py:1651-1682, 2039-2054
# --------------------------------------------------------------------------------------------


def minibatch_update(X, sample_weight, centers, counts, labels):
    """In-place running-mean update.  Sums run over samples in index order, products and sums
    rounded separately (the Cython is compiled without FMA contraction on x86-64 baseline)."""
    K = centers.shape[0]
    for k in range(K):
        idx = np.where(labels == k)[0]
        wsum = 0.0
        for i in idx:
            wsum += sample_weight[i]
        if wsum > 0:
            acc = centers[k] * counts[k]
            for i in idx:
                acc = acc + X[i] * sample_weight[i]
            counts[k] += wsum
            alpha = 1.0 / counts[k]
            centers[k] = acc * alpha
    return centers, counts


def mini_batch_step(X, sample_weight, centers, counts, random_state, random_reassign, reassignment_ratio=0.01):
    """One ``_mini_batch_step`` (sklearn/cluster/_kmeans.py:1566-1684), dense, in place."""
    labels = kmeans_assign_tiebreak(X, centers)
    minibatch_update(X, sample_weight, centers, counts, labels)
    if random_reassign and reassignment_ratio > 0:
        to_reassign = counts < reassignment_ratio * counts.max()
        if to_reassign.sum() > 0.5 * X.shape[0]:
            keep = np.argsort(counts)[int(0.5 * X.shape[0]):]
            to_reassign[keep] = False
        n_reassigns = to_reassign.sum()
        if n_reassigns:
            new_centers = random_state.choice(X.shape[0], replace=False, size=n_reassigns)
            centers[to_reassign] = X[new_centers]
        counts[to_reassign] = np.min(counts[~to_reassign])
    return labels


# --------------------------------------------------------------------------------------------
# a5'': one full Lloyd iteration   sklearn/cluster/_k_means_lloyd.pyx:23-165 (dense, 1 thread)
# --------------------------------------------------------------------------------------------


def lloyd_iter(X, sample_weight, centers):
    """Returns (labels, new_centers, weight_in_clusters).  Empty clusters keep their old centre
    here (sklearn relocates them to the farthest points, _k_means_common.pyx; the GPU path leaves
    that rare host-side decision to the caller and the tests avoid it)."""
    labels = kmeans_assign(X, centers)
    K, D = centers.shape
    sums = np.zeros((K, D))
    wsum = np.zeros(K)
    for i in range(X.shape[0]):
        k = labels[i]
        wsum[k] += sample_weight[i]
        sums[k] = sums[k] + X[i] * sample_weight[i]
    new = centers.copy()
    nz = wsum > 0
    new[nz] = sums[nz] * (1.0 / wsum[nz])[:, None]
    return labels, new, wsum


def lloyd_iter_fast(X, sample_weight, centers):
    """Vectorised Lloyd iteration (np.add.at keeps index order) for larger cases."""
    labels = kmeans_assign(X, centers)
    K, D = centers.shape
    sums = np.zeros((K, D))
    wsum = np.zeros(K)
    np.add.at(wsum, labels, sample_weight)
    np.add.at(sums, labels, X * sample_weight[:, None])
    new = centers.copy()
    nz = wsum > 0
    new[nz] = sums[nz] * (1.0 / wsum[nz])[:, None]
    return labels, new, wsum


# --------------------------------------------------------------------------------------------
# a6/a7: flux matrix      reference: msm_we/_hamsm/_fluxmatrix.py:21-72, 97-164, 166-345
# --------------------------------------------------------------------------------------------


def build_flux_matrix(n_clusters, index_pairs, ind_start_in_basis, ind_end_in_basis, ind_end_in_target, transition_weights):
    """_fluxmatrix.py:97-164 -- relabel in the reference's order, then coo_matrix."""
    basis = n_clusters
    target = n_clusters + 1
    start, end = np.asarray(index_pairs).T.copy()
    end[ind_end_in_target] = target
    start[ind_start_in_basis] = basis
    end[ind_end_in_basis] = basis
    return coo_matrix((transition_weights, (start, end)), shape=(n_clusters + 2, n_clusters + 2))


def iter_flux_matrix(n_clusters, index_pairs, pcoord0, pcoord1, weights, basis_bounds, target_bounds):
    """_fluxmatrix.py:21-72 -- dense per-iteration matrix."""
    end_t = np.where(is_we_region(pcoord1, target_bounds))
    start_b = np.where(is_we_region(pcoord0, basis_bounds))
    end_b = np.where(is_we_region(pcoord1, basis_bounds))
    return np.asarray(build_flux_matrix(n_clusters, index_pairs, start_b, end_b, end_t, weights).todense())


def flux_matrix(n_clusters, per_iter, basis_bounds, target_bounds):
    """_fluxmatrix.py:232-260, 342 -- serial path: sum dense per-iteration matrices in order, / nI.

    ``per_iter``: iterable of (index_pairs[S,2], pcoord0[S,P], pcoord1[S,P], weights[S]).
    """
    M = n_clusters + 2
    total = np.zeros((M, M))
    nI = 0
    for pairs, p0, p1, w in per_iter:
        total = total + iter_flux_matrix(n_clusters, pairs, p0, p1, w, basis_bounds, target_bounds)
        nI += 1
    return total / nI


# --------------------------------------------------------------------------------------------
# a9: colour-augmented count scatter       reference: msm_we/nmm.py:117-167
# --------------------------------------------------------------------------------------------


def colour_counts(trajectories, n_states, state_a, state_b, lag, sliding_window=True):
    """2N x 2N history-coloured count matrix, row = 2*s_prev + colour_prev, col = 2*s_now + colour_now
    (A -> 0, B -> 1); transitions with an undefined previous colour are skipped (nmm.py:132-158)."""
    nm = np.zeros((2 * n_states, 2 * n_states))
    A, B = set(state_a), set(state_b)
    step = 1 if sliding_window else lag
    for traj in trajectories:
        traj = np.asarray(traj)
        for start in range(lag, 2 * lag, step):
            prev = -1
            for i in range(start, len(traj), lag):
                s = int(traj[i])
                col = 0 if s in A else (1 if s in B else prev)
                if prev >= 0 and col >= 0:
                    nm[2 * int(traj[i - lag]) + prev, 2 * s + col] += 1.0
                prev = col
    return nm


def colour_transitions(trajectories, state_a, state_b, lag, sliding_window=True):
    """The same walk, but emitting the (state_prev, state_now, colour_prev, colour_now) records
    that feed the GPU flux kernel with C=2."""
    A, B = set(state_a), set(state_b)
    step = 1 if sliding_window else lag
    s0, s1, c0, c1 = [], [], [], []
    for traj in trajectories:
        traj = np.asarray(traj)
        for start in range(lag, 2 * lag, step):
            prev = -1
            for i in range(start, len(traj), lag):
                s = int(traj[i])
                col = 0 if s in A else (1 if s in B else prev)
                if prev >= 0 and col >= 0:
                    s0.append(int(traj[i - lag])); s1.append(s); c0.append(prev); c1.append(col)
                prev = col
    return (np.array(s0, dtype=np.int64), np.array(s1, dtype=np.int64),
            np.array(c0, dtype=np.uint8), np.array(c1, dtype=np.uint8))


def normalize_markov_matrix(m):
    """Row-normalise, zero rows stay zero (msm_we/utils.py normalize_markov_matrix, non-reversible)."""
    m = np.array(m, dtype=np.float64)
    rs = m.sum(axis=1)
    nz = rs != 0
    m[nz] = m[nz] / rs[nz, None]
    return m


# --------------------------------------------------------------------------------------------
# a3/a4/a5: per-iteration discretization and streaming stratified clustering
#           reference: msm_we/_hamsm/_clustering.py:748-918, 1144-1329
# --------------------------------------------------------------------------------------------


def discretize_iteration(strat: StratifiedOracle, Xp, Xc, pcoord0, pcoord1, literal=False, models=None):
    """_clustering.py:1298-1316: parents binned by pcoord0List, children by pcoord1List."""
    parent = strat.predict(Xp, pcoord0, literal=literal, models=models)
    child = strat.predict(Xc, pcoord1, literal=literal, models=models)
    return parent, child


def stratified_clustering_batches(data, iters, bin_mapper, n_clusters_per_bin, basis_bounds, target_bounds,
                                  use_weights=False):
    """The batching rule of do_stratified_clustering (_clustering.py:794-916), as a generator of
    ``(used_iters, [(bin, X_bin, w_bin_or_None), ...])``.

    ``data.iteration(it)`` must return an object with ``pcoord0 [S,P]``, ``child [S,D]`` (already
    featurised+transformed) and ``weights [S]``.

    Reproduces the reference's index handling to the letter: basis/target parents are dropped from
    the pcoord array only (_clustering.py:872-874), and the per-bin row indices computed on that
    filtered array are applied to the *unfiltered* coordinates and weights (:892-899).
    """
    iters = list(iters)
    pos = 0
    while pos < len(iters):
        used = -1
        coords = None
        weights = None
        pcoords = []
        unique_bins = np.array([])
        counts = np.array([])
        assignments = np.array([])
        filled = False
        while not filled:
            if pos + used + 1 >= len(iters):
                # out of iterations: unfilled bins are folded into the nearest filled bin (:806-826)
                unfilled = unique_bins[counts < n_clusters_per_bin]
                filled_bins = np.setdiff1d(unique_bins, unfilled)
                for ub in unfilled:
                    nearest = find_nearest_bin(bin_mapper, int(ub), [int(f) for f in filled_bins])
                    # (reference indexes with a tuple from np.where; same effect)
                    assignments[assignments == ub] = nearest
                unique_bins = filled_bins
                break
            used += 1
            it = data.iteration(iters[pos + used])
            if used == 0:
                coords = it.child
                weights = it.weights
                pcoords = [p for p in it.pcoord0]
            else:
                coords = np.append(coords, it.child, axis=0)
                pcoords.extend(it.pcoord0)
                if use_weights:
                    weights = np.append(weights, it.weights, axis=0)
            parr = np.array(pcoords)
            drop = is_we_region(parr, target_bounds) | is_we_region(parr, basis_bounds)
            parr = parr[~drop]
            assignments = bin_mapper.assign(parr) if len(parr) > 0 else np.array([])
            unique_bins, counts = np.unique(assignments, return_counts=True)
            filled = bool(np.all(counts >= n_clusters_per_bin))
        batch = []
        for b in unique_bins:
            rows = np.where(assignments == b)[0]
            Xb = coords[rows]
            wb = weights[rows] if use_weights else None
            batch.append((int(b), Xb, wb))
        yield used, batch
        pos += 1 + max(used, 0)


def find_nearest_bin(bin_mapper, bin_idx, filled_bins):
    """_clustering.py:1331-1396, rectilinear / Voronoi-Euclidean."""
    assert len(filled_bins) > 0
    if hasattr(bin_mapper, "centers"):
        centers = np.asarray(bin_mapper.centers, dtype=np.float64)
    else:
        mids = [np.asarray(b[:-1]) + (np.asarray(b[1:]) - np.asarray(b[:-1])) / 2 for b in bin_mapper.boundaries]
        centers = np.array(np.meshgrid(*mids)).T.squeeze().reshape(-1, len(mids))
    ignored = np.setdiff1d(range(centers.shape[0]), filled_bins)
    others = np.delete(centers, ignored, axis=0)
    with np.errstate(invalid="ignore"):
        d = np.sqrt(np.mean(np.power(centers[bin_idx] - others, 2), axis=1))
    closest = int(np.argmin(d))
    for ib in sorted(ignored):
        if closest >= ib:
            closest += 1
    return closest


# --------------------------------------------------------------------------------------------
# downstream check: transition matrix + steady state     msm_we/_hamsm/_analysis.py:23-79, 97-191
# (reported, not a parity gate)
# --------------------------------------------------------------------------------------------


def transition_matrix(flux, ind_basis, ind_targets):
    f = np.array(flux, dtype=np.float64)
    out = f.sum(axis=1)
    for s in range(f.shape[0]):
        if out[s] > 0:
            f[s, :] = f[s, :] / out[s]
        if out[s] == 0.0:
            f[s, s] = 1.0
    sink = np.zeros((1, f.shape[0]))
    sink[0, ind_basis] = 1.0 / np.size(ind_basis)
    f[ind_targets, :] = np.tile(sink, (np.size(ind_targets), 1))
    return f


def steady_state(tmatrix):
    """Left eigenvector for the eigenvalue closest to 1, normalised to sum 1."""
    vals, vecs = np.linalg.eig(tmatrix.T)
    k = np.argmin(np.abs(vals - 1.0))
    p = np.real(vecs[:, k])
    p = p / p.sum()
    return p


def linear_transform(X, components, mean=None):
    """reference: ``self.coordinates.transform`` of the fitted IncrementalPCA (msm_we/_hamsm/_dimensionality.py:243;
    sklearn < 1.1 ``_BasePCA.transform``: ``X = X - self.mean_; np.dot(X, self.components_.T)``), applied right
    before every predict (msm_we/_hamsm/_clustering.py:1291-1296)."""
    X = np.asarray(X, dtype=np.float64)
    if mean is not None:
        X = X - np.asarray(mean, dtype=np.float64)
    return np.dot(X, np.asarray(components, dtype=np.float64).T)
