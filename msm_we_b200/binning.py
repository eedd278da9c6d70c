"""WE bin mappers with the westpa interface the hot path uses (``nbins``, ``assign``, ``boundaries`` /
``centers``).

The reference takes a ``westpa.core.binning`` mapper (msm_we/_hamsm/_clustering.py:588-609 accepts only
``RectilinearBinMapper`` and ``VoronoiBinMapper``); westpa is not importable in this image, so these two
classes restate the part of that interface msm_we touches.  A real westpa mapper of either type is
accepted too (duck-typed on ``boundaries`` / ``centers``); anything else is used through its own host
``assign`` and the bins are handed to the GPU as precomputed.

``assign`` here is host code for the *control* decisions of the clustering driver (a few thousand
pcoords per batch); the per-frame bin lookup of the discretization runs in the K0 kernel.
"""
from __future__ import annotations

import numpy as np


class RectilinearBinMapper:
    """float32 boundaries, ``b[i] <= x < b[i+1]`` per dimension, row-major bin index (last dimension
    fastest), ``ValueError`` for a coordinate outside the bin space."""

    def __init__(self, boundaries):
        self._boundaries = [np.asarray(b, dtype=np.float32) for b in boundaries]
        for b in self._boundaries:
            if b.ndim != 1 or len(b) < 2 or not np.all(np.diff(b) > 0):
                raise ValueError("boundaries must be 1-D, strictly increasing, at least 2 per dimension")
        self.ndim = len(self._boundaries)
        self.nbins = int(np.prod([len(b) - 1 for b in self._boundaries]))
        self.labels = [f"bin {i}" for i in range(self.nbins)]

    @property
    def boundaries(self):
        return self._boundaries

    def assign(self, coords, mask=None, output=None):
        coords = np.asarray(coords)
        if coords.ndim == 1:
            coords = coords[:, None]
        c32 = coords.astype(np.float32)
        index = np.zeros(c32.shape[0], dtype=np.int64)
        for d, b in enumerate(self._boundaries):
            pos = np.searchsorted(b, c32[:, d], side="right") - 1
            if ((pos < 0) | (pos >= len(b) - 1) | np.isnan(c32[:, d])).any():
                raise ValueError("coordinate outside of bin space")
            index = index * (len(b) - 1) + pos
        if output is not None:
            output[...] = index
            return output
        return index


class VoronoiBinMapper:
    """Nearest-centre bins.  ``dfunc(coord, centers) -> distances``; the default (and the only one the
    GPU evaluates natively) is the squared Euclidean distance in float32."""

    def __init__(self, dfunc=None, centers=None, dfargs=None, dfkwargs=None):
        if centers is None:
            raise ValueError("centers are required")
        self.centers = np.asarray(centers, dtype=np.float32)
        if self.centers.ndim == 1:
            self.centers = self.centers[:, None]
        self.ndim = self.centers.shape[1]
        self.nbins = self.centers.shape[0]
        self.dfunc = dfunc if dfunc is not None else self._euclidean
        self.is_euclidean = dfunc is None
        self.dfargs = dfargs or ()
        self.dfkwargs = dfkwargs or {}
        self.labels = [f"center={c!r}" for c in self.centers]

    @staticmethod
    def _euclidean(coord, centers, *args, **kwargs):
        diff = np.asarray(centers, dtype=np.float32) - np.asarray(coord, dtype=np.float32)
        return np.sqrt(np.mean(diff * diff, axis=-1))

    def assign(self, coords, mask=None, output=None):
        coords = np.asarray(coords)
        if coords.ndim == 1:
            coords = coords[:, None]
        c32 = coords.astype(np.float32)
        if self.is_euclidean:
            d2 = np.zeros((c32.shape[0], self.nbins), dtype=np.float32)
            for d in range(self.ndim):
                diff = c32[:, d:d + 1] - self.centers[None, :, d]
                d2 = d2 + diff * diff
            index = np.argmin(d2, axis=1).astype(np.int64)
        else:
            index = np.array([int(np.argmin(self.dfunc(c, self.centers, *self.dfargs, **self.dfkwargs))) for c in c32],
                             dtype=np.int64)
        if output is not None:
            output[...] = index
            return output
        return index


SUPPORTED_MAPPERS = {RectilinearBinMapper, VoronoiBinMapper}


def mapper_kind(bin_mapper) -> str:
    """'rectilinear' / 'voronoi' when the GPU can evaluate the mapper itself, else 'host'."""
    name = type(bin_mapper).__name__
    if isinstance(bin_mapper, RectilinearBinMapper) or (name == "RectilinearBinMapper" and hasattr(bin_mapper, "boundaries")):
        return "rectilinear"
    if isinstance(bin_mapper, VoronoiBinMapper):
        return "voronoi" if bin_mapper.is_euclidean else "host"
    return "host"
