"""Host-side analysis steps that consume the cleaned flux matrix: transition matrix, steady state, target flux.

reference: msm_we/_hamsm/_analysis.py -- ``get_Tmatrix`` (:23-79), ``get_steady_state`` (:97-191),
``get_steady_state_algebraic`` (:193-282), ``get_steady_state_target_flux`` (:317-384) and the helpers
``inverse_iteration`` / ``is_connected`` (msm_we/utils.py:87-161).

This is NOT part of the GPU hot path (SURVEY section 8 keeps this small dense / sparse linear algebra on the host); it is
here so that ``build_analyze_model``, ``do_block_validation`` and the ``HAMSMDriver`` plugin run to the end -- ``pSS`` and
``JtargetSS`` are what their callers read -- without mixing the reference's own AnalysisMixin in.  numpy / scipy on
(nBins x nBins) matrices, same results as the reference run (tests/test_reference_fixtures.py: ``Tmatrix`` 1e-10, ``pSS``
and ``JtargetSS`` 1e-6).
"""
from __future__ import annotations

import numpy as np

from .._logging import log


def _reaches(tmatrix, sources, targets):
    """Every target is reachable from every source along non-zero entries (utils.py:87-113)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import shortest_path

    dist = shortest_path(csr_matrix(tmatrix), directed=True, indices=np.atleast_1d(sources))
    return not np.isinf(dist[:, np.atleast_1d(targets)]).any()


def _inverse_iteration_step(tmatrix_sparse, guess, mu=1.0):
    """One step of inverse iteration for the left eigenvector of eigenvalue ``mu``: solve ``(T^T - mu I) y = guess`` and
    normalise to sum 1 (utils.py:116-161; the reference forms the explicit sparse inverse, a factorisation solves the
    same system).  A factorisation that reports an exactly singular matrix is retried with ``mu = 0.999``, as there."""
    from scipy.sparse import identity
    from scipy.sparse.linalg import splu

    n = guess.shape[0]
    system = (tmatrix_sparse.T - mu * identity(n, format="csc")).tocsc()
    try:
        y = splu(system).solve(np.asarray(guess, dtype=np.float64))
        if not np.isfinite(y).all():
            raise RuntimeError("inverse iteration produced non-finite values")
    except RuntimeError:
        if mu != 1.0:
            raise
        log.error("inverse iteration failed at mu = 1, retrying with mu = 0.999")
        return _inverse_iteration_step(tmatrix_sparse, guess, mu=0.999)
    y = np.asarray(y).squeeze()
    return y / y.sum()


class AnalysisMixin:
    Tmatrix = None
    pSS = None
    JtargetSS = None
    lagtime = None

    def get_Tmatrix(self):
        """Row-normalised flux matrix; a state without outgoing flux keeps itself; every target state recycles uniformly
        into the basis states (reference :23-79).  Sets ``self.Tmatrix``."""
        flux = np.array(self.fluxMatrix, dtype=np.float64)
        out = flux.sum(axis=1)
        t = flux.copy()
        moving = out > 0
        t[moving] = flux[moving] / out[moving][:, None]
        stuck = np.flatnonzero(out == 0.0)
        t[stuck, stuck] = 1.0
        recycle = np.zeros(self.nBins)
        recycle[self.indBasis] = 1.0 / np.size(self.indBasis)
        t[np.atleast_1d(self.indTargets), :] = recycle
        self.Tmatrix = t

    def get_steady_state_algebraic(self, max_iters=1000, check_negative=True, set=True):
        """Left eigenvector of the largest (real part) eigenvalue from the dense eigensolver, normalised; when it carries
        negative entries, up to ``max_iters`` applications of growing powers of the transition matrix try to remove them
        (reference :193-282)."""
        vals, vecs = np.linalg.eig(self.Tmatrix.T)
        pss = np.real(vecs[:, np.argmax(np.real(vals))]).squeeze()
        assert not np.isclose(pss.sum(), 0), "Steady-state distribution sums to 0!"
        pss = pss / pss.sum()
        if (pss < 0).any() and max_iters > 0:
            last, power, fixed = pss, self.Tmatrix.copy(), None
            for _ in range(max_iters):
                new = power.T @ last
                if not (new < 0).any():
                    fixed = new
                    break
                last = new
                power = self.Tmatrix @ power
            if fixed is None:
                log.warning("Power method did NOT obtain semidefinite pSS. Some negative values remain.")
            else:
                pss = fixed
        if not (pss >= 0).all():
            if check_negative:
                raise AssertionError(f"Some negative elements in steady-state distribution: {pss}")
            log.warning("Some negative elements in pSS... Ignoring, and setting model.pSS anyways.")
        if set:
            self.pSS = pss
            return None
        return pss

    def get_steady_state(self, flux_fractional_convergence=1e-4, max_iters=10):
        """Eigensolver estimate refined by inverse iteration on the sparse matrix until the target flux changes by less
        than ``flux_fractional_convergence`` of its value (reference :97-191).  Sets ``self.pSS``."""
        from scipy.sparse import csr_matrix

        sparse_t = csr_matrix(self.Tmatrix)
        pss = self.get_steady_state_algebraic(max_iters=10, check_negative=False, set=False)
        flux = self.get_steady_state_target_flux(pSS=pss, _set=False)
        for it in range(max_iters):
            pss = _inverse_iteration_step(sparse_t, pss)
            new_flux = self.get_steady_state_target_flux(pSS=pss, _set=False)
            change, flux = new_flux - flux, new_flux
            if abs(change) < flux * flux_fractional_convergence:
                log.info(f"Flux converged to {flux:.4e} after {it + 1} iterations of inverse iteration.")
                break
            if it == max_iters - 1 and flux != 0:
                log.warning("Flux is nonzero and did not converge!")
        assert (pss >= 0).all(), "Negative elements in pSS"
        assert flux >= 0, "Negative flux estimate from this pSS"
        self.pSS = pss

    def get_steady_state_target_flux(self, pSS=None, _set=True):
        """Probability flux per unit time into the target states: ``sum_{i not target} pSS_i T_ij`` over the targets ``j``,
        divided by the lag time ``tau (n_lag + 1)`` (reference :317-384).  Returns -1 when no path leads from the basis to
        the target."""
        if not _reaches(self.Tmatrix, self.indBasis, self.indTargets):
            log.critical("There is no path in this matrix from the basis to the target, so no MFPT can be calculated.")
            return -1
        pss = np.squeeze(np.asarray(self.pSS if pSS is None else pSS))
        lagtime = self.tau * (self.n_lag + 1)
        targets = np.atleast_1d(self.indTargets)
        others = np.setdiff1d(np.arange(self.nBins), targets)
        total = 0.0
        for j in targets:
            total = total + np.sum(pss[others] * self.Tmatrix[others, j])
        if _set:
            self.lagtime = lagtime
            self.JtargetSS = total / lagtime
            return None
        return total / lagtime
