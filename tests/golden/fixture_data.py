"""Seeded WE data sets behind the reference-executed fixtures (tests/golden/make_reference_fixtures.py).

TEST INFRASTRUCTURE ONLY.  The small data sets are stored inside the fixtures; the larger ones are regenerated
from their seed in the tests and checked against a checksum stored in the fixture, so a test can never silently
compare against results for different inputs.
"""
from __future__ import annotations

import hashlib

import numpy as np


def checksum(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def we_dataset(seed, n_iters, segs0, seg_growth, n_atoms, coord_ndim, bins_per_dim, k_true, pcoord_ndim=1, step=0.7,
               noise=1.0, skip_bin=None, skip_until=None, regions_last=None):
    """A weighted-ensemble-like run: every segment's parent is a random segment of the previous iteration (its
    start pcoord / coordinates ARE that segment's end pcoord / coordinates), the child takes a reflected Gaussian
    step in pcoord space, coordinates are one of ``k_true`` micro-state means of the WE bin + noise, weights span
    many decades and sum to one per iteration.  Returns a list over iterations of dicts in the layout of
    ``refshim.register_we_file``.

    ``regions_last=(basis_bounds, target_bounds)``: segments whose PARENT pcoord lies in the basis or target are
    placed last in their iteration.  do_stratified_clustering drops those rows from the pcoord array but not from the
    coordinate array (_clustering.py:872-899), which scrambles the bin <-> coordinate pairing of
    every row after the first dropped one; with the dropped rows last, single-iteration batches stay aligned and the
    fitted centres are meaningful.  Without it the fixture exercises the scrambled pairing.

    ``skip_bin``/``skip_until``: pcoords that would land in WE bin ``skip_bin`` (dimension 0) before iteration
    ``skip_until`` are pushed one bin up, so that bin is never seen while clustering on the early iterations."""
    rng = np.random.default_rng(seed)
    D = n_atoms * coord_ndim
    nb = bins_per_dim
    nbins = nb ** pcoord_ndim
    means = rng.normal(0, 3, size=(nbins, 1, D)) + rng.normal(0, 2, size=(nbins, k_true, D))
    hi = nb - 1e-3

    def flat_bin(pc):
        idx = np.zeros(len(pc), dtype=np.int64)
        for d in range(pcoord_ndim):
            idx = idx * nb + np.minimum(pc[:, d].astype(np.int64), nb - 1)
        return idx

    def feats(pc):
        k = rng.integers(0, k_true, size=len(pc))
        return means[flat_bin(pc), k] + rng.normal(0, noise, size=(len(pc), D))

    def fix(pc, it):
        if skip_bin is not None and it < skip_until:
            sel = (pc[:, 0] >= skip_bin) & (pc[:, 0] < skip_bin + 1)
            pc[sel, 0] += 1.0
        return pc

    prev_pc = fix(rng.uniform(0, hi, size=(segs0, pcoord_ndim)), 0)
    prev_x = feats(prev_pc)
    its = []
    for i in range(n_iters):
        S = segs0 + seg_growth * i
        par = rng.integers(0, len(prev_pc), size=S)
        pc0, xp = prev_pc[par], prev_x[par]
        pc1 = np.abs(pc0 + rng.normal(0, step, size=(S, pcoord_ndim)))
        pc1 = np.where(pc1 > hi, 2 * hi - pc1, pc1).clip(0, hi)
        pc1 = fix(pc1, i + 1)
        if regions_last is not None:
            drop = np.zeros(S, dtype=bool)
            for bounds in regions_last:
                inside = np.ones(S, dtype=bool)
                for d, (lo, hi_) in enumerate(bounds):
                    inside &= (pc0[:, d] > lo) & (pc0[:, d] < hi_)
                drop |= inside
            order = np.argsort(drop, kind="stable")
            par, pc0, xp, pc1 = par[order], pc0[order], xp[order], pc1[order]
        xc = feats(pc1)
        w = np.exp(rng.normal(0, 3, size=S))
        w /= w.sum()
        its.append(dict(weights=w, pcoord=np.stack([pc0, pc1], axis=1),
                        coords=np.stack([xp, xc], axis=1).reshape(S, 2, n_atoms, coord_ndim), parent_id=par))
        prev_pc, prev_x = pc1, xc
    return its


def boundaries(bins_per_dim, pcoord_ndim=1):
    b = np.arange(bins_per_dim + 1, dtype=np.float32)
    b[-1] = np.inf
    return [b.copy() for _ in range(pcoord_ndim)]


def pack_iterations(its):
    """Flat arrays for np.savez."""
    lens = np.array([len(it["weights"]) for it in its], dtype=np.int64)
    return dict(in_lens=lens, in_weights=np.concatenate([it["weights"] for it in its]),
                in_pcoord=np.concatenate([it["pcoord"] for it in its]),
                in_coords=np.concatenate([it["coords"] for it in its]),
                in_parent_id=np.concatenate([it["parent_id"] for it in its]))


def unpack_iterations(d):
    offs = np.concatenate([[0], np.cumsum(d["in_lens"])])
    return [dict(weights=d["in_weights"][a:b], pcoord=d["in_pcoord"][a:b], coords=d["in_coords"][a:b],
                 parent_id=d["in_parent_id"][a:b]) for a, b in zip(offs[:-1], offs[1:])]


def flatten_featurizer(self, coords):
    """A user featuriser as the HAMSMDriver plugin loads one (``featurization: fixture_data.flatten_featurizer``)."""
    coords = np.asarray(coords)
    if coords.ndim == 2:
        return coords.reshape(1, -1)
    return coords.reshape(coords.shape[0], -1)
