"""Iteration feeder: the fields of ``modelWE`` the hot path reads, served from an iteration source.

reference: msm_we/_hamsm/_data.py -- ``load_iter_data`` (:807-932), ``get_transition_data_lag0``
(:254-320), ``load_iter_coordinates`` (:557-618), ``get_iter_coordinates`` (:531-555), ``get_coordSet``
(:677-759), ``get_iterations`` (:934-993); on-disk layout: msm_we/westpa_plugins/augmentation_driver.py:173-180
(``iterations/iter_%08d/{seg_index, pcoord [S, pcoord_len, P], auxdata/<auxpath> [S, 2, nAtoms, 3]}``).

The reference opens every file once per iteration and grows ``pcoord0List`` / ``weightList`` with one
``np.append`` per segment.  Here an *iteration source* hands out whole iterations:

* ``H5IterationSource`` -- WESTPA HDF5 files (needs ``h5py``): the iteration groups of every file are indexed
  once, every field is ONE bulk dataset read, and ``read_pair_into`` lets the GPU staging code read the
  start / end structures of an iteration straight into its pinned host buffer (``Dataset.read_direct``), so
  the coordinates cross host memory once on their way to the device;
* ``ArrayIterationSource`` -- in-memory arrays (tests, synthetic benchmarks, users who already hold their
  trajectories in numpy).

What the methods set is exactly what the reference sets: ``n_iter``, ``nSeg``, ``weightList``, ``westList``,
``segindList``, ``pcoord0List``, ``pcoord1List``, ``seg_weights[n_iter]``, ``coordPairList [nSeg, nAtoms,
coord_ndim, 2]``, ``transitionWeights``, ``departureWeights``, ``cur_iter_coords``, ``numSegments``, ``maxIter``,
``pcoordSet``.
"""
from __future__ import annotations

import collections

import numpy as np

from .._logging import log


class IterationRecord:
    """One WE iteration: start/end pcoords, weights and start/end coordinates of every segment (plus, when the
    source knows them, the WE parent of every segment and which file / row it came from)."""

    __slots__ = ("pcoord0", "pcoord1", "weights", "parent_coords", "child_coords", "parent_id", "west_file", "seg_index")

    def __init__(self, pcoord0, pcoord1, weights, parent_coords, child_coords, parent_id=None, west_file=None,
                 seg_index=None):
        self.pcoord0 = np.asarray(pcoord0, dtype=np.float64)
        self.pcoord1 = np.asarray(pcoord1, dtype=np.float64)
        if self.pcoord0.ndim == 1:
            self.pcoord0 = self.pcoord0[:, None]
        if self.pcoord1.ndim == 1:
            self.pcoord1 = self.pcoord1[:, None]
        self.weights = np.asarray(weights, dtype=np.float64)
        self.parent_coords = None if parent_coords is None else np.asarray(parent_coords, dtype=np.float64)
        self.child_coords = None if child_coords is None else np.asarray(child_coords, dtype=np.float64)
        n = self.weights.shape[0]
        self.parent_id = None if parent_id is None else np.asarray(parent_id, dtype=np.int64)
        self.west_file = np.zeros(n, dtype=np.int64) if west_file is None else np.asarray(west_file, dtype=np.int64)
        self.seg_index = np.arange(n, dtype=np.int64) if seg_index is None else np.asarray(seg_index, dtype=np.int64)
        lens = [self.pcoord0.shape[0], self.pcoord1.shape[0]]
        lens += [a.shape[0] for a in (self.parent_coords, self.child_coords) if a is not None]
        if any(k != n for k in lens):
            raise ValueError("all per-segment arrays of an iteration must have the same length")


class ArrayIterationSource:
    """Iterations 1..n held in memory.  Coordinates may be ``[S, nAtoms, 3]`` structures or ``[S, F]``
    feature rows (treated as ``nAtoms=F, coord_ndim=1``).  The source OWNS its arrays for the lifetime of the
    model, which is what allows the staging code to page-lock them in place."""

    owns_arrays = True

    def __init__(self, records=None):
        self._records = {}
        for i, r in enumerate(records or [], start=1):
            self._records[i] = r

    def add(self, n_iter, record: IterationRecord):
        self._records[int(n_iter)] = record

    def has(self, n_iter):
        return int(n_iter) in self._records

    def get(self, n_iter) -> IterationRecord:
        return self._records[int(n_iter)]

    def n_segments(self, n_iter):
        return self._records[int(n_iter)].weights.shape[0]

    def n_iterations(self):
        n = 0
        while (n + 1) in self._records:
            n += 1
        return n

    # a model is deep-copied by the reference's flow (pre_discretization_model, post_cluster_model, one copy per
    # block-validation group); the records are read-only inputs, so copies share them instead of duplicating the
    # whole data set in host RAM
    def __deepcopy__(self, memo):
        return self

    def __copy__(self):
        return self


class H5IterationSource:
    """WESTPA ``west.h5`` reader.  An iteration counts only when the NEXT iteration's ``seg_index`` exists in the
    same file (the last iteration of a run holds no dynamics; reference _data.py:866-869, 968-972); segments of one
    iteration may be spread over several files and are concatenated in file order, as the reference does."""

    owns_arrays = False
    CACHE = 4

    def __init__(self, file_list, auxpath="coord", pcoord_ndim=1, pcoord_len=2):
        try:
            import h5py  # noqa: F401
        except ImportError as e:
            raise ImportError("reading WESTPA HDF5 files needs h5py; pass an ArrayIterationSource instead") from e
        self.file_list = list(file_list)
        self.auxpath = auxpath
        self.pcoord_ndim = int(pcoord_ndim)
        self.pcoord_len = int(pcoord_len)
        self._index = None
        self._cache = collections.OrderedDict()
        self.pcoord_shape_warned = False

    # open handles are not picklable / copyable state: only the description travels
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_cache"] = collections.OrderedDict()
        return state

    def __deepcopy__(self, memo):
        return self

    @staticmethod
    def _group(n_iter):
        return "iterations/iter_%08d" % int(n_iter)

    def _open(self, name):
        import h5py

        return h5py.File(name, "r")

    def _build_index(self):
        """iteration -> [(file index, n segments)], from ONE pass over the group names of every file."""
        index = {}
        for fi, name in enumerate(self.file_list):
            f = self._open(name)
            try:
                if "iterations" not in f:
                    continue
                names = set(f["iterations"].keys())
                for g in sorted(names):
                    if not g.startswith("iter_"):
                        continue
                    n = int(g[5:])
                    grp = "iterations/" + g
                    if f"{grp}/seg_index" in f and ("iter_%08d" % (n + 1)) in names \
                            and f"{self._group(n + 1)}/seg_index" in f:
                        index.setdefault(n, []).append((fi, int(f[f"{grp}/seg_index"].shape[0])))
            finally:
                f.close()
        self._index = index

    def _files(self, n_iter):
        if self._index is None:
            self._build_index()
        return self._index.get(int(n_iter), [])

    def has(self, n_iter):
        return len(self._files(n_iter)) > 0

    def n_segments(self, n_iter):
        return sum(s for _, s in self._files(n_iter))

    def n_iterations(self):
        n = 0
        while self.has(n + 1):
            n += 1
        return n

    def get(self, n_iter, coords=True) -> IterationRecord:
        """All fields of one iteration with one bulk read per dataset.  ``coords=False`` skips the structures
        (``load_iter_data`` needs only pcoords and weights)."""
        key = (int(n_iter), bool(coords))
        if key in self._cache:
            self._cache.move_to_end(key)
            return self._cache[key]
        if coords is False and (int(n_iter), True) in self._cache:
            return self._cache[(int(n_iter), True)]
        files = self._files(n_iter)
        if not files:
            raise KeyError(f"iteration {n_iter} is in none of {self.file_list}")
        P, L = self.pcoord_ndim, self.pcoord_len
        p0, p1, w, par, pc, cc, wf, si = [], [], [], [], [], [], [], []
        for fi, _ in files:
            f = self._open(self.file_list[fi])
            try:
                grp = self._group(n_iter)
                seg = f[f"{grp}/seg_index"][:]
                pcoord = f[f"{grp}/pcoord"][:]
                if pcoord.shape[2] != P and not self.pcoord_shape_warned:
                    log.warning(f"Dimensions of pcoord in {self.file_list[fi]} ({pcoord.shape[2]}) do not match specified "
                                f"pcoord dimensionality self.pcoord_ndim ({P}). MSM-WE will only load up to dimension {P}.")
                    self.pcoord_shape_warned = True
                S = seg.shape[0]
                w.append(np.asarray(seg["weight"], dtype=np.float64))
                par.append(np.asarray(seg["parent_id"], dtype=np.int64) if "parent_id" in (seg.dtype.names or ())
                           else np.full(S, -1, dtype=np.int64))
                p0.append(pcoord[:, 0, :P])
                p1.append(pcoord[:, L - 1, :P])
                wf.append(np.full(S, fi, dtype=np.int64))
                si.append(np.arange(S, dtype=np.int64))
                if coords:
                    dset = f[f"{grp}/auxdata/{self.auxpath}"]     # KeyError when the run was not augmented
                    if dset.shape[1] <= 1:
                        raise AssertionError("Augmented coords only have 1 point in them -- need at least start & end "
                                             "for transitions")
                    c = dset[:]
                    pc.append(c[:, 0])
                    cc.append(c[:, L - 1])
            finally:
                f.close()
        cat = np.concatenate
        rec = IterationRecord(cat(p0), cat(p1), cat(w), cat(pc) if coords else None, cat(cc) if coords else None,
                              parent_id=cat(par), west_file=cat(wf), seg_index=cat(si))
        self._cache[key] = rec
        while len(self._cache) > self.CACHE:
            self._cache.popitem(last=False)
        return rec

    def read_pair_into(self, n_iter, dst_parent, dst_child):
        """Start / end structures of every segment, flattened to ``[S, nAtoms*coord_ndim]`` rows, read DIRECTLY
        into the caller's (pinned) buffers: ``Dataset.read_direct`` with a source selection, no intermediate array.
        Returns False when the iteration cannot be served that way (then the caller goes through ``get``)."""
        files = self._files(n_iter)
        pos = 0
        L = self.pcoord_len
        for fi, S in files:
            f = self._open(self.file_list[fi])
            try:
                dset = f[f"{self._group(n_iter)}/auxdata/{self.auxpath}"]
                if not hasattr(dset, "read_direct") or dset.dtype != np.float64:
                    return False
                shape = (S,) + tuple(dset.shape[2:])
                for dst, t in ((dst_parent, 0), (dst_child, L - 1)):
                    view = dst[pos:pos + S].reshape(shape)
                    dset.read_direct(view, np.s_[:, t], np.s_[...])
            finally:
                f.close()
            pos += S
        return True


class DataMixin:
    n_iter = None
    fileList = None
    n_data_files = None
    numSegments = None
    maxIter = None
    weightList = None
    westList = None
    segindList = None
    nSeg = None
    pcoord0List = None
    pcoord1List = None
    seg_weights = {}
    coordPairList = None
    transitionWeights = None
    departureWeights = None
    coordsExist = None
    pcoordSet = None
    iteration_source = None

    def _record(self, n_iter, coords=True) -> IterationRecord:
        if self.iteration_source is None:
            raise RuntimeError("model has no iteration source; call initialize() first")
        src = self.iteration_source
        if isinstance(src, H5IterationSource):
            return src.get(n_iter, coords=coords)
        return src.get(n_iter)

    @staticmethod
    def _as_structures(coords):
        # feature rows [S, F] are carried as [S, F, 1] so the reference's (nSeg, nAtoms, coord_ndim) shapes hold
        return coords[:, :, None] if coords.ndim == 2 else coords

    def load_iter_data(self, n_iter: int):
        """reference: _data.py:807-932."""
        self.n_iter = n_iter
        if not self.iteration_source.has(n_iter):
            self.weightList = np.array([])
            self.westList = np.array([], dtype=int)
            self.segindList = np.array([], dtype=int)
            self.nSeg = 0
            self.pcoord0List = np.empty((0, self.pcoord_ndim))
            self.pcoord1List = np.empty((0, self.pcoord_ndim))
            self.seg_weights[n_iter] = np.array([])
            return
        rec = self._record(n_iter, coords=False)
        self.seg_weights[n_iter] = rec.weights.copy()
        self.weightList = rec.weights.copy()
        self.westList = rec.west_file.astype(int)
        self.segindList = rec.seg_index.astype(int)
        self.nSeg = rec.weights.shape[0]
        self.pcoord0List = rec.pcoord0[:, : self.pcoord_ndim].copy()
        self.pcoord1List = rec.pcoord1[:, : self.pcoord_ndim].copy()

    def get_transition_data_lag0(self):
        """reference: _data.py:254-320 (segments with NaN coordinates get weight 0)."""
        weightList = self.weightList
        if self.nSeg == 0:
            self.coordPairList = np.zeros((0, self.nAtoms or 0, self.coord_ndim or 3, 2))
            self.transitionWeights = weightList.copy()
            self.departureWeights = weightList.copy()
            return
        rec = self._record(self.n_iter)
        parent = self._as_structures(rec.parent_coords)
        child = self._as_structures(rec.child_coords)
        coordPairList = np.zeros((self.nSeg, parent.shape[1], parent.shape[2], 2))
        coordPairList[:, :, :, 0] = parent
        coordPairList[:, :, :, 1] = child
        nan_segments = np.where(np.isnan(coordPairList).any(axis=(1, 2, 3)))[0]
        if nan_segments.shape[0] > 0:
            log.warning(f"Bad coordinates for segments {nan_segments}, setting weights to 0")
            weightList[nan_segments] = 0.0
        self.coordPairList = coordPairList
        self.transitionWeights = weightList.copy()
        self.departureWeights = weightList.copy()

    # ---- bulk accessors used by the batched GPU paths (no [nSeg, nAtoms, 3, 2] intermediate) ----------
    def iter_coordinate_pair(self, n_iter):
        """(parent, child) structures ``[S, nAtoms, coord_ndim]`` of one iteration, straight from the source."""
        rec = self._record(n_iter)
        return self._as_structures(rec.parent_coords), self._as_structures(rec.child_coords)

    def iter_nan_segments(self, n_iter):
        """Indices of segments with a NaN start or end coordinate (the reference zeroes their weights,
        _data.py:302-313).  Cached per iteration: the discretization pass touches the coordinates anyway."""
        cache = self.__dict__.setdefault("_nan_segments_cache", {})
        if n_iter not in cache:
            parent, child = self.iter_coordinate_pair(n_iter)
            S = parent.shape[0]
            # one pass, no temporaries: a row sum is NaN whenever the row holds a NaN; the (rare) candidates
            # are then checked exactly, so inf - inf cannot produce a false positive
            cand = np.where(np.isnan(parent.reshape(S, -1).sum(axis=1)) | np.isnan(child.reshape(S, -1).sum(axis=1)))[0]
            if cand.shape[0]:
                exact = np.isnan(parent[cand]).any(axis=(1, 2)) | np.isnan(child[cand]).any(axis=(1, 2))
                cand = cand[exact]
            cache[n_iter] = cand
        return cache[n_iter]

    def note_nan_segments(self, n_iter, rows):
        """Lets a pass that already has the coordinates on the device (discretization) record the NaN segments, so
        the flux pass does not have to re-read the structures just for the weights."""
        self.__dict__.setdefault("_nan_segments_cache", {})[n_iter] = np.asarray(rows, dtype=np.int64)

    def iter_transition_weights(self, n_iter):
        """``transitionWeights`` of get_transition_data_lag0 without materialising coordPairList."""
        w = self._record(n_iter, coords=False).weights.copy()
        bad = self.iter_nan_segments(n_iter)
        if bad.shape[0] > 0:
            w[bad] = 0.0
        return w

    def load_iter_coordinates(self):
        """reference: _data.py:557-618 (end-of-segment coordinates of the loaded iteration; NaN rows and
        ``coordsExist = False`` when the iteration holds no augmented coordinates)."""
        if self.nSeg == 0:
            self.cur_iter_coords = np.full((0, self.nAtoms or 0, self.coord_ndim or 3), fill_value=np.nan)
            return
        try:
            self.cur_iter_coords = self._as_structures(self._record(self.n_iter).child_coords).copy()
        except KeyError:
            log.error(f"Error getting coordinates in iteration {self.n_iter}")
            self.cur_iter_coords = np.full((self.nSeg, self.nAtoms or 0, self.coord_ndim or 3), fill_value=np.nan)
            self.coordsExist = False

    def get_iter_coordinates(self, iteration):
        """reference: _data.py:531-555 (rows with NaN coordinates are dropped)."""
        self.load_iter_data(iteration)
        self.load_iter_coordinates()
        bad = np.isnan(self.cur_iter_coords).any(axis=(1, 2))
        return self.cur_iter_coords[~bad]

    def get_iterations(self):
        """reference: _data.py:934-993."""
        src = self.iteration_source
        n = src.n_iterations()
        self.numSegments = np.array([float(src.n_segments(i)) for i in range(1, n + 1)])
        self.maxIter = self.numSegments.size

    def get_coordSet(self, last_iter, streaming=None, progress_bar=None):
        """reference: _data.py:677-759.  Always streams (as the reference effectively does): only ``pcoordSet``
        -- the end pcoord of every segment of iterations 1..last_iter, NaN rows where the end structure is bad --
        is held; the coordinates themselves stay in the iteration source."""
        total = int(sum(self.numSegments[:last_iter]))
        pcoordSet = np.full((total, self.pcoord_ndim), fill_value=np.nan)
        pos = 0
        for i in range(1, last_iter + 1):
            if not self.iteration_source.has(i):
                continue
            rec = self._record(i, coords=False)
            S = rec.weights.shape[0]
            pcoordSet[pos:pos + S] = rec.pcoord1[:, : self.pcoord_ndim]
            bad = self._bad_end_structures(i)
            if bad.shape[0]:
                log.warning(f"Walker {bad} has bad coords in iteration(s) {i}")
                pcoordSet[pos + bad] = np.nan
            pos += S
        self.pcoordSet = pcoordSet
        self.first_iter = 1
        self.last_iter = last_iter
        # the reference walks the iterations downwards and leaves the model on iteration 1
        if last_iter >= 1:
            self.load_iter_data(1)
            self.load_iter_coordinates()

    def _bad_end_structures(self, n_iter):
        """Rows whose END structure holds a NaN (get_coordSet / get_iter_coordinates semantics)."""
        cache = self.__dict__.setdefault("_bad_end_cache", {})
        if n_iter not in cache:
            try:
                child = self._as_structures(self._record(n_iter).child_coords)
            except KeyError:
                cache[n_iter] = np.arange(self.iteration_source.n_segments(n_iter))
                return cache[n_iter]
            S = child.shape[0]
            cand = np.where(np.isnan(child.reshape(S, -1).sum(axis=1)))[0]
            if cand.shape[0]:
                cand = cand[np.isnan(child[cand]).any(axis=(1, 2))]
            cache[n_iter] = cand
        return cache[n_iter]
