"""Iteration feeder: the fields of ``modelWE`` the hot path reads, served from an iteration source.

reference: msm_we/_hamsm/_data.py -- ``load_iter_data`` (:807-932), ``get_transition_data_lag0``
(:254-320), ``load_iter_coordinates`` (:557-618), ``get_iter_coordinates`` (:531-555),
``get_iterations`` (:934-993).  The reference reads WESTPA HDF5 files with h5py (per-segment
``np.append`` loops); HDF5 ingestion is outside this build's scope and h5py is
not installed here, so the same attributes are filled from an *iteration source*:

* ``ArrayIterationSource`` -- in-memory arrays (tests, synthetic benchmarks, users who already hold
  their trajectories in numpy);
* ``H5IterationSource`` -- bulk reads of ``iterations/iter_%08d/{seg_index,pcoord,auxdata/<auxpath>}``
  when h5py is importable (layout written by msm_we/westpa_plugins/augmentation_driver.py:173-180).

What the methods set is exactly what the reference sets: ``n_iter``, ``nSeg``, ``weightList``,
``pcoord0List``, ``pcoord1List``, ``seg_weights[n_iter]``, ``coordPairList [nSeg, nAtoms, coord_ndim, 2]``,
``transitionWeights``, ``departureWeights``, ``cur_iter_coords``, ``numSegments``, ``maxIter``.
"""
from __future__ import annotations

import numpy as np

from .._logging import log


class IterationRecord:
    """One WE iteration: start/end pcoords, weights and start/end coordinates of every segment."""

    __slots__ = ("pcoord0", "pcoord1", "weights", "parent_coords", "child_coords")

    def __init__(self, pcoord0, pcoord1, weights, parent_coords, child_coords):
        self.pcoord0 = np.asarray(pcoord0, dtype=np.float64)
        self.pcoord1 = np.asarray(pcoord1, dtype=np.float64)
        if self.pcoord0.ndim == 1:
            self.pcoord0 = self.pcoord0[:, None]
        if self.pcoord1.ndim == 1:
            self.pcoord1 = self.pcoord1[:, None]
        self.weights = np.asarray(weights, dtype=np.float64)
        self.parent_coords = np.asarray(parent_coords, dtype=np.float64)
        self.child_coords = np.asarray(child_coords, dtype=np.float64)
        n = self.weights.shape[0]
        if not (self.pcoord0.shape[0] == self.pcoord1.shape[0] == self.parent_coords.shape[0]
                == self.child_coords.shape[0] == n):
            raise ValueError("all per-segment arrays of an iteration must have the same length")


class ArrayIterationSource:
    """Iterations 1..n held in memory.  Coordinates may be ``[S, nAtoms, 3]`` structures or ``[S, F]``
    feature rows (treated as ``nAtoms=F, coord_ndim=1``)."""

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

    def n_iterations(self):
        n = 0
        while (n + 1) in self._records:
            n += 1
        return n


class H5IterationSource:
    """WESTPA west.h5 reader (needs h5py).  Bulk dataset reads; the last iteration of a file is treated
    as incomplete, as in the reference (_data.py:876-879)."""

    def __init__(self, file_list, auxpath="coord", pcoord_ndim=1):
        try:
            import h5py  # noqa: F401
        except ImportError as e:  # pragma: no cover - h5py is absent in the build image
            raise ImportError("reading WESTPA HDF5 files needs h5py; pass an ArrayIterationSource instead") from e
        self.file_list = list(file_list)
        self.auxpath = auxpath
        self.pcoord_ndim = pcoord_ndim

    def _open(self, n_iter):  # pragma: no cover
        import h5py

        for name in self.file_list:
            f = h5py.File(name, "r")
            if f"/iterations/iter_{int(n_iter):08d}/seg_index" in f and \
                    f"/iterations/iter_{int(n_iter) + 1:08d}/seg_index" in f:
                yield f
            f.close()

    def has(self, n_iter):  # pragma: no cover
        return any(True for _ in self._open(n_iter))

    def get(self, n_iter):  # pragma: no cover
        p0, p1, w, pc, cc = [], [], [], [], []
        for f in self._open(n_iter):
            grp = f[f"/iterations/iter_{int(n_iter):08d}"]
            pcoord = grp["pcoord"][:]
            p0.append(pcoord[:, 0, : self.pcoord_ndim])
            p1.append(pcoord[:, -1, : self.pcoord_ndim])
            w.append(grp["seg_index"]["weight"])
            coords = grp[f"auxdata/{self.auxpath}"]
            pc.append(coords[:, 0])
            cc.append(coords[:, -1])
        return IterationRecord(np.concatenate(p0), np.concatenate(p1), np.concatenate(w), np.concatenate(pc),
                               np.concatenate(cc))

    def n_iterations(self):  # pragma: no cover
        n = 0
        while self.has(n + 1):
            n += 1
        return n


class DataMixin:
    n_iter = None
    fileList = None
    n_data_files = None
    numSegments = None
    maxIter = None
    weightList = None
    nSeg = None
    pcoord0List = None
    pcoord1List = None
    seg_weights = {}
    coordPairList = None
    transitionWeights = None
    departureWeights = None
    coordsExist = None
    iteration_source = None

    def _record(self, n_iter) -> IterationRecord:
        if self.iteration_source is None:
            raise RuntimeError("model has no iteration source; call initialize() first")
        return self.iteration_source.get(n_iter)

    @staticmethod
    def _as_structures(coords):
        # feature rows [S, F] are carried as [S, F, 1] so the reference's (nSeg, nAtoms, coord_ndim) shapes hold
        return coords[:, :, None] if coords.ndim == 2 else coords

    def load_iter_data(self, n_iter: int):
        """reference: _data.py:807-932."""
        self.n_iter = n_iter
        if not self.iteration_source.has(n_iter):
            self.weightList = np.array([])
            self.nSeg = 0
            self.pcoord0List = np.empty((0, self.pcoord_ndim))
            self.pcoord1List = np.empty((0, self.pcoord_ndim))
            self.seg_weights[n_iter] = np.array([])
            return
        rec = self._record(n_iter)
        self.seg_weights[n_iter] = rec.weights.copy()
        self.weightList = rec.weights.copy()
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

    def iter_transition_weights(self, n_iter):
        """``transitionWeights`` of get_transition_data_lag0 without materialising coordPairList."""
        w = self._record(n_iter).weights.copy()
        bad = self.iter_nan_segments(n_iter)
        if bad.shape[0] > 0:
            w[bad] = 0.0
        return w

    def load_iter_coordinates(self):
        """reference: _data.py:557-618 (end-of-segment coordinates of the loaded iteration)."""
        if self.nSeg == 0:
            self.cur_iter_coords = np.full((0, self.nAtoms or 0, self.coord_ndim or 3), fill_value=np.nan)
            return
        self.cur_iter_coords = self._as_structures(self._record(self.n_iter).child_coords).copy()

    def get_iter_coordinates(self, iteration):
        """reference: _data.py:531-555 (rows with NaN coordinates are dropped)."""
        self.load_iter_data(iteration)
        self.load_iter_coordinates()
        bad = np.isnan(self.cur_iter_coords).any(axis=(1, 2))
        return self.cur_iter_coords[~bad]

    def get_iterations(self):
        """reference: _data.py:934-993."""
        n = self.iteration_source.n_iterations()
        self.numSegments = np.array([float(self.iteration_source.get(i).weights.shape[0]) for i in range(1, n + 1)])
        self.maxIter = self.numSegments.size
