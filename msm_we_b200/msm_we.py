"""``modelWE``: the reference's model-builder class, restricted to the hot path this package accelerates.

reference: msm_we/msm_we.py -- class composition (:35-42), ``initialize`` (:143-277), bounds setters
(:279-440), ``is_WE_basis`` / ``is_WE_target`` (:462-527), ``build_analyze_model`` (:588-882).

Kept: every name the discretization + flux path reads or writes.  Not re-implemented (out
of scope, host-side small-matrix work in the reference): dimensionality-reduction fitting, flux-matrix
cleaning, transition matrix / steady state / committors, block validation, plotting.  Those mixins of
the reference can be combined with this class unchanged because the attributes they consume
(``fluxMatrixRaw``, ``dtrajs``, ``pair_dtrajs``, ``clusters`` ...) have the reference's types.
"""
from __future__ import annotations

import numpy as np

from ._hamsm._analysis import AnalysisMixin
from ._hamsm._clustering import ClusteringMixin
from ._hamsm._data import ArrayIterationSource, DataMixin, H5IterationSource
from ._hamsm._fluxmatrix import FluxMatrixMixin
from ._logging import log


class Coordinates:
    """Identity transform (reference: msm_we/_hamsm/_dimensionality.py:23-34)."""

    is_identity = True

    def __init__(self):
        self.explanation = "coordinate object"

    def transform(self, coords):
        return coords


class LinearCoordinates:
    """Fitted affine projection ``(X - mean_) @ components_.T`` -- what the reference's IncrementalPCA
    ``coordinates.transform`` applies immediately before assignment (_dimensionality.py:243)."""

    def __init__(self, components, mean=None):
        self.components_ = np.asarray(components, dtype=np.float64)
        self.mean_ = np.zeros(self.components_.shape[1]) if mean is None else np.asarray(mean, dtype=np.float64)

    def transform(self, coords):
        return (np.asarray(coords, dtype=np.float64) - self.mean_) @ self.components_.T

    def device_projection(self, device):
        """(components [d_out, D_in], mean [D_in]) as CUDA tensors: ``launch_ray_discretization`` then ships the
        featurised frames as they are and applies the projection on the device (``ops.project``), instead of a
        host matmul per iteration."""
        import torch

        cache = self.__dict__.setdefault("_device_cache", {})
        key = str(device)
        if key not in cache:
            cache[key] = (torch.from_numpy(np.ascontiguousarray(self.components_)).to(device),
                          torch.from_numpy(np.ascontiguousarray(self.mean_)).to(device))
        return cache[key]

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_device_cache", None)
        return state


class modelWE(ClusteringMixin, DataMixin, FluxMatrixMixin, AnalysisMixin):
    class BlockValidationError(Exception):
        pass

    def __init__(self):
        self.modelName = None
        self._n_lag = 0
        self.pcoord_ndim = None
        self.pcoord_len = None
        self.tau = None
        self._basis_pcoord_bounds = None
        self._target_pcoord_bounds = None
        self.nAtoms = None
        self.coord_ndim = 3
        self.coordinates = None
        self.ndim = None
        self.removed_clusters = []
        self.cluster_structures = None
        self.cluster_structure_weights = None
        self.clustering_method = None
        self.validation_models = []
        self.pcoord_shape_warned = False
        self.pre_discretization_model = None
        self.dimReduceMethod = "none"
        self.seg_weights = {}
        # (sic) the reference initialises these singular names to None and never sets them; get_cluster_centers reads
        # them (msm_we.py:98,109; _clustering.py:1544-1555)
        self.target_bin_center = None
        self.basis_bin_center = None
        self.reference_coord = None
        self.reference_structure = None
        self.basis_coords = None
        self.indBasis = None
        self.indTargets = None

    # the reference expects users to monkey-patch this featurizer (docs/usage.rst:41-60)
    def processCoordinates(self, coords):
        coords = np.asarray(coords)
        if coords.ndim == 3:
            return coords.reshape(coords.shape[0], -1)
        if coords.ndim == 2 and self.nAtoms is not None and coords.shape == (self.nAtoms, self.coord_ndim):
            return coords.reshape(1, -1)
        return coords

    # marks the default featuriser: when nothing but this flatten stands between a stored structure and its feature
    # row, the staging code may read structures from the iteration source straight into its pinned buffer
    processCoordinates._mwe_flatten = True

    # ------------------------------------------------------------------------------------------
    def initialize(self, fileSpecifier, refPDBfile=None, modelName=None, basis_pcoord_bounds=None,
                   target_pcoord_bounds=None, dim_reduce_method="none", tau=None, pcoord_ndim=1, auxpath="coord",
                   _suppress_boundary_warning=False, use_weights_in_clustering=False):
        """reference: msm_we.py:143-277.  ``fileSpecifier`` is a list of WESTPA HDF5 paths (needs h5py) or an
        iteration source object (``ArrayIterationSource``)."""
        log.debug("Initializing msm_we model")
        self.modelName = modelName
        self.pcoord_ndim = pcoord_ndim
        self.pcoord_len = 2
        if basis_pcoord_bounds is None:
            log.warning("No basis coord bounds provided to initialize().")
        else:
            self.basis_pcoord_bounds = basis_pcoord_bounds
        if target_pcoord_bounds is None:
            log.warning("No target coord bounds provided to initialize().")
        else:
            self.target_pcoord_bounds = target_pcoord_bounds
        self.auxpath = auxpath
        if hasattr(fileSpecifier, "get") and hasattr(fileSpecifier, "has"):
            self.iteration_source = fileSpecifier
            self.fileList = []
        else:
            if isinstance(fileSpecifier, str):
                log.warning("HDF5 file paths were provided in a string -- this is now deprecated, please pass as a list "
                            "of paths.")
                files = fileSpecifier.split(" ")
            else:
                files = list(fileSpecifier)
            self.fileList = files
            self.iteration_source = H5IterationSource(files, auxpath=auxpath, pcoord_ndim=pcoord_ndim,
                                                      pcoord_len=self.pcoord_len)
        self.n_data_files = len(self.fileList)
        if tau is None:
            log.warning("No tau provided, defaulting to 1.")
            tau = 1.0
        self.tau = float(tau)
        if refPDBfile is not None:
            self.set_topology(refPDBfile)
        if dim_reduce_method is None:
            log.warning("No dimensionality reduction method provided to initialize(). Defaulting to pca.")
            self.dimReduceMethod = "pca"
        else:
            self.dimReduceMethod = dim_reduce_method
        if self.dimReduceMethod == "none":
            self.coordinates = Coordinates()
        self.use_weights_in_clustering = use_weights_in_clustering
        try:
            self.load_iter_data(1)
            rec = self._record(1)
            c = rec.child_coords
            if self.nAtoms is None:
                self.nAtoms = c.shape[1]
                self.coord_ndim = c.shape[2] if c.ndim == 3 else 1
            self.coordsExist = True
        except KeyError:
            if not _suppress_boundary_warning:
                log.warning("Model initialized, but coordinates do not exist yet.")
            self.coordsExist = False
        log.debug("msm_we model successfully initialized")

    def set_topology(self, topology):
        """reference: msm_we.py:1011-1078.  Only ``nAtoms`` / ``coord_ndim`` matter to the hot path; a dict
        ``{"coords", "nAtoms", "coord_ndim"}`` needs no MD library, PDB / prmtop paths go through mdtraj when it is
        importable."""
        if isinstance(topology, dict):
            self.reference_coord = topology["coords"]
            self.nAtoms = topology["nAtoms"]
            self.coord_ndim = topology["coord_ndim"]
            return
        if isinstance(topology, str) and topology[-3:] == "dat":
            self.reference_coord = np.loadtxt(topology)
            self.nAtoms = 1
            self.coord_ndim = 3
            return
        try:
            import mdtraj as md
        except ImportError:
            log.warning("mdtraj is not importable: the topology is not loaded, nAtoms is taken from the stored coordinates")
            return
        if isinstance(topology, str):
            struct = md.load_prmtop(topology) if topology[-6:] == "prmtop" else md.load(topology)
        elif type(topology) in [md.Trajectory, md.Topology]:
            struct = topology
        else:
            raise NotImplementedError("Unsupported topology")
        self.reference_structure = struct
        self.nAtoms = struct.n_atoms if hasattr(struct, "n_atoms") and not hasattr(struct, "topology") else struct.topology.n_atoms
        if hasattr(struct, "_xyz"):
            self.reference_coord = np.squeeze(struct._xyz)
        self.coord_ndim = 3

    def dimReduce(self, *args, **kwargs):
        """reference: _dimensionality.py:110-345.  FITTING a dimensionality reduction (streaming PCA / TICA / VAMP) is
        outside the path this package covers; APPLYING a fitted one is on it (``LinearCoordinates`` runs on the
        device).  So: ``"none"`` installs the identity; any other method requires that a fitted object with a
        ``.transform`` has been assigned to ``self.coordinates`` beforehand -- silently clustering on raw features
        instead would give a different discretization than the reference's."""
        if self.dimReduceMethod == "none":
            self.coordinates = Coordinates()
            self.ndim = None if self.nAtoms is None else int(self.coord_ndim) * int(self.nAtoms)
            return
        if self.coordinates is None or isinstance(self.coordinates, Coordinates):
            raise NotImplementedError(
                f"dim_reduce_method={self.dimReduceMethod!r}: msm_we_b200 does not fit dimensionality reductions. Fit it "
                f"with the reference (or sklearn) and assign the result to model.coordinates (e.g. "
                f"LinearCoordinates(components_, mean_)) before dimReduce(); arguments {sorted(kwargs)} were not used.")
        comp = getattr(self.coordinates, "components_", None)
        if comp is not None:
            self.ndim = int(np.shape(comp)[0])

    # ---- bounds (reference: msm_we.py:279-440) -------------------------------------------------
    def _check_bounds(self, bounds):
        bounds = np.array(bounds)
        if len(bounds.shape) == 1:
            log.warning("Please provide 1-D boundaries as a list of lists or 2-D array [[lower bound, upper bound]]. "
                        "Automatically doing conversion for now.")
            bounds = np.reshape(bounds, (1, 2))
        assert bounds.shape == (self.pcoord_ndim, 2), \
            f"Shape of bounds was {bounds.shape}, should've been ({self.pcoord_ndim}, 2)"
        assert np.all([b[0] < b[1] for b in bounds]), "A boundary has a lower bound larger than its upper bound"
        centers = np.full(self.pcoord_ndim, fill_value=np.nan)
        for i, (lo, hi) in enumerate(bounds):
            if not abs(lo) == np.inf and not abs(hi) == np.inf:
                centers[i] = np.mean([lo, hi])
            else:
                centers[i] = [lo, hi][abs(lo) == np.inf]
        return bounds, centers

    @property
    def basis_pcoord_bounds(self):
        return self._basis_pcoord_bounds

    @basis_pcoord_bounds.setter
    def basis_pcoord_bounds(self, bounds):
        self._basis_pcoord_bounds, self.basis_bin_centers = self._check_bounds(bounds)

    @property
    def target_pcoord_bounds(self):
        return self._target_pcoord_bounds

    @target_pcoord_bounds.setter
    def target_pcoord_bounds(self, bounds):
        self._target_pcoord_bounds, self.target_bin_centers = self._check_bounds(bounds)

    @property
    def WEbasisp1_bounds(self):
        return self.basis_pcoord_bounds

    @WEbasisp1_bounds.setter
    def WEbasisp1_bounds(self, bounds):
        self.basis_pcoord_bounds = bounds

    @property
    def WEtargetp1_bounds(self):
        return self.target_pcoord_bounds

    @WEtargetp1_bounds.setter
    def WEtargetp1_bounds(self, bounds):
        if None in bounds:
            raise Exception("A target boundary has not been correctly provided")
        self.target_pcoord_bounds = bounds

    @property
    def n_lag(self):
        return self._n_lag

    @n_lag.setter
    def n_lag(self, lag):
        if not lag == 0:
            raise NotImplementedError("Only a lag of 1 tau (n_lag = 0) is currently supported")
        self._n_lag = lag

    # ---- region tests (reference: msm_we.py:462-527); host versions for the control flow ----------
    def _in_region(self, pcoords, bounds):
        pcoords = np.asarray(pcoords)
        inside = np.full_like(pcoords, fill_value=np.nan, dtype=np.float64)
        for d in range(self.pcoord_ndim):
            inside[:, d] = np.logical_and(pcoords[:, d] > bounds[d, 0], pcoords[:, d] < bounds[d, 1])
        return np.all(inside, axis=1)

    def is_WE_basis(self, pcoords):
        return self._in_region(pcoords, self.basis_pcoord_bounds)

    def is_WE_target(self, pcoords):
        return self._in_region(pcoords, self.target_pcoord_bounds)

    # ---- one-shot driver (reference: msm_we.py:588-882), hot-path steps ----------------------------
    HOST_ANALYSIS_STEPS = ("get_Tmatrix", "get_steady_state", "get_steady_state_target_flux")

    def _host_analysis(self):
        """Transition matrix, steady state and target flux of the cleaned flux matrix: small host-side linear algebra
        (``_hamsm/_analysis.py``; the north star keeps it on the host), in the reference's order (msm_we.py:812-841)."""
        for step in self.HOST_ANALYSIS_STEPS:
            getattr(self, step)()
        return True

    def build_analyze_model(self, file_paths, ref_struct, modelName, basis_pcoord_bounds, target_pcoord_bounds,
                            dimreduce_method, tau, n_clusters, ray_kwargs={}, max_coord_iter=-1, stratified=True,
                            streaming=True, use_ray=True, fluxmatrix_iters=[1, -1], fluxmatrix_iters_to_use=None,
                            cross_validation_groups=2, cross_validation_blocks=4, show_live_display=True,
                            allow_validation_failure=False, step_kwargs={}):
        """reference: msm_we.py:588-882, same steps in the same order (no Ray initialisation, no live table)."""
        model = self
        model.initialize(fileSpecifier=file_paths, refPDBfile=ref_struct, modelName=modelName,
                         basis_pcoord_bounds=basis_pcoord_bounds, target_pcoord_bounds=target_pcoord_bounds,
                         dim_reduce_method=dimreduce_method, tau=tau, **step_kwargs.get("initialize", {}))
        model.get_iterations()
        _max_coord_iter = [max_coord_iter, model.maxIter][max_coord_iter == -1]
        model.get_coordSet(_max_coord_iter)
        model.dimReduce(**step_kwargs.get("dimReduce", {}))
        model.cluster_coordinates(n_clusters=n_clusters, streaming=streaming, use_ray=use_ray, stratified=stratified,
                                  store_validation_model=cross_validation_groups > 0,
                                  **step_kwargs.get("clustering", {}))
        # the reference mutates its default list here; work on a copy
        _fluxmatrix_iters = list(fluxmatrix_iters)
        if _fluxmatrix_iters[1] == -1:
            _fluxmatrix_iters[1] = model.maxIter
        model.get_fluxMatrix(n_lag=0, first_iter=_fluxmatrix_iters[0], last_iter=_fluxmatrix_iters[1],
                             iters_to_use=fluxmatrix_iters_to_use, use_ray=use_ray, **step_kwargs.get("fluxmatrix", {}))
        model.organize_fluxMatrix(use_ray=use_ray, **step_kwargs.get("organize", {}))
        model._host_analysis()
        if cross_validation_groups > 0:
            try:
                model.do_block_validation(cross_validation_groups=cross_validation_groups,
                                          cross_validation_blocks=cross_validation_blocks, use_ray=use_ray,
                                          **step_kwargs.get("block_validation", {}))
            except Exception as e:
                log.error(e)
                if not allow_validation_failure:
                    raise e
        return model

    def do_block_validation(self, cross_validation_groups, cross_validation_blocks, use_ray=True, progress_bar=None):
        """reference: msm_we.py:884-1009.  The iterations are cut into ``cross_validation_blocks`` uniform blocks dealt
        round-robin to ``cross_validation_groups`` groups; every group gets a copy of the post-clustering model and runs
        the flux hot path (K3 over its iteration subset), the cleaning pass (re-discretize + re-flux) and, when mixed in,
        the host analysis.  Copies share the iteration source, so a group costs the model state, not the data set."""
        from copy import deepcopy

        assert hasattr(self, "post_cluster_model") and self.post_cluster_model is not None, (
            "Perform clustering with cluster_coordinates() before attempting"
            "block validation -- self.post_cluster_model is not set.")
        validation_models = [deepcopy(self.post_cluster_model) for _ in range(cross_validation_groups)]
        max_iter = self.post_cluster_model.maxIter
        iters_per_block = max_iter // cross_validation_blocks
        block_iterations = [[start, start + iters_per_block] for start in range(1, max_iter, iters_per_block)]
        block_iterations[-1][-1] = block_iterations[-1][-1] - 1
        group_blocks = [range(start, cross_validation_blocks, cross_validation_groups)
                        for start in range(cross_validation_groups)]
        validation_iterations = []
        for group in range(cross_validation_groups):
            group_iterations = []
            for block in group_blocks[group]:
                group_iterations.extend(range(*block_iterations[block]))
            validation_iterations.append(group_iterations)
            try:
                log.info(f"Beginning analysis of cross-validation group {group + 1}/{cross_validation_groups}.")
                _model = validation_models[group]
                _model.get_fluxMatrix(0, iters_to_use=validation_iterations[group], use_ray=use_ray,
                                      progress_bar=progress_bar)
                _model.organize_fluxMatrix(use_ray=use_ray, progress_bar=progress_bar)
                _model._host_analysis()
            except Exception as e:
                log.error("Error during block validation!")
                log.exception(e)
                raise modelWE.BlockValidationError(e)
        self.validation_iterations = validation_iterations
        self.validation_models = validation_models


__all__ = ["modelWE", "Coordinates", "LinearCoordinates", "ArrayIterationSource"]
