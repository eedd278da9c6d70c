"""WESTPA plugin boundary: ``HAMSMDriver`` with the reference's configuration keys, building the haMSM
through the GPU hot path.

reference: msm_we/westpa_plugins/hamsm_driver.py:8-144.  WESTPA itself is optional here (it is not
installed in the build image): status messages go through ``westpa.rc.pstatus`` when westpa imports and
through the package logger otherwise, and the featurizer is resolved with ``westpa.core.extloader`` or
``importlib``.  Everything else -- callback registration on ``sim_manager.finalize_run``, the plugin
keys, the ``build_analyze_model`` call and ``data_manager.hamsm_model`` -- is as in the reference.
"""
from __future__ import annotations

import importlib

from .. import msm_we
from .._logging import log


def _pstatus(msg):
    try:
        import westpa

        westpa.rc.pstatus(msg)
    except Exception:
        log.info(msg)


def _get_object(name):
    try:
        from westpa.core import extloader

        return extloader.get_object(name)
    except ImportError:
        module, _, attr = name.rpartition(".")
        return getattr(importlib.import_module(module), attr)


class HAMSMDriver:
    def __init__(self, sim_manager, plugin_config):
        _pstatus("Initializing haMSM plugin")
        if not sim_manager.work_manager.is_master:
            _pstatus("Not running on the master process, skipping")
            return
        self.data_manager = sim_manager.data_manager
        self.sim_manager = sim_manager
        self.plugin_config = plugin_config
        self.priority = plugin_config.get("priority", 2)
        sim_manager.register_callback(sim_manager.finalize_run, self.construct_hamsm, self.priority)
        self.h5file_paths = [self.data_manager.we_h5filename]
        self.first_iter_to_use = self.plugin_config.get("first_analysis_iter", 1)
        self.dimreduce_use_weights = self.plugin_config.get("dimreduce_use_weights", True)
        self.dimreduce_var_cutoff = self.plugin_config.get("dimreduce_var_cutoff", None)
        self.cross_validation_groups = self.plugin_config.get("cross_validation_groups", 2)
        self.ray_address = self.plugin_config.get("ray_address", None)   # accepted, unused on the GPU path
        self.ray_kwargs = self.plugin_config.get("ray_kwargs", {})

    def construct_hamsm(self):
        self.data_manager.hamsm_model = None
        refPDBfile = self.plugin_config.get("ref_pdb_file")
        model_name = self.plugin_config.get("model_name")
        clusters_per_stratum = self.plugin_config.get("n_clusters")
        target_pcoord_bounds = self.plugin_config.get("target_pcoord_bounds")
        basis_pcoord_bounds = self.plugin_config.get("basis_pcoord_bounds")
        dimreduce_method = self.plugin_config.get("dimreduce_method", None)
        tau = self.plugin_config.get("tau", None)

        featurization_module = self.plugin_config.get("featurization")
        featurizer = _get_object(featurization_module)
        msm_we.modelWE.processCoordinates = featurizer
        self.data_manager.processCoordinates = featurizer
        self.data_manager.close_backing()

        ray_kwargs = {"num_cpus": self.plugin_config.get("num_cpus", None)}
        ray_kwargs.update(self.ray_kwargs)
        clustering_kwargs = {"first_cluster_iter": self.first_iter_to_use}
        # optional: a user bin mapper object / cluster arguments can ride in the plugin config
        if self.plugin_config.get("user_bin_mapper") is not None:
            clustering_kwargs["user_bin_mapper"] = self.plugin_config.get("user_bin_mapper")
        clustering_kwargs.update(self.plugin_config.get("cluster_args", {}))

        model = msm_we.modelWE()
        model.build_analyze_model(
            file_paths=self.h5file_paths, ref_struct=refPDBfile, modelName=model_name,
            basis_pcoord_bounds=basis_pcoord_bounds, target_pcoord_bounds=target_pcoord_bounds,
            dimreduce_method=dimreduce_method, n_clusters=clusters_per_stratum, tau=tau, ray_kwargs=ray_kwargs,
            step_kwargs={
                "dimReduce": {"use_weights": self.dimreduce_use_weights, "variance_cutoff": self.dimreduce_var_cutoff,
                              "first_iter": self.first_iter_to_use, "first_rough_iter": self.first_iter_to_use},
                "clustering": clustering_kwargs,
            },
            fluxmatrix_iters=[self.first_iter_to_use, -1], allow_validation_failure=True,
            cross_validation_groups=self.cross_validation_groups)
        _pstatus(f"Storing built haMSM on {self.data_manager}")
        self.data_manager.hamsm_model = model
        return model
