"""Generate the committed golden fixtures from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py

Outputs (all under tests/golden/):
  ntl9_clustered.npz   from tests/reference/1000ns_ntl9/models/clustered.obj (stub-unpickled):
                       per-bin cluster centres, bin boundaries, we_remap, pair_dtrajs (parent, child
                       labels per iteration), basis/target pcoord bounds, n_clusters; and the
                       non-zero pattern of models/fluxmatrix_raw.npy.
  ntl9_downstream.npz  models/fluxmatrix.npy, tmatrix.npy, pSS.npy, JtargetSS.npy and the basis /
                       target state indices of organized.obj (the downstream steady-state check).
  colour_kat.npz       the known-answer 6x6 matrices printed in the reference's
                       tests/test_non_markov_model.py:15-24 and test_markov_color_model.py:16-25.
The pickles are read with a stub unpickler (msm_we / mdtraj / westpa / sklearn classes become
attribute bags), so none of those packages is needed.
"""
import io
import os
import pickle
import sys

import numpy as np

REF = "/root/reference/tests/reference/1000ns_ntl9/models"
OUT = os.path.dirname(os.path.abspath(__file__))


class _Bag:
    def __new__(cls, *args, **kwargs):
        return object.__new__(cls)

    def __init__(self, *args, **kwargs):
        if args:
            self.__dict__["_args"] = args

    def __call__(self, *args, **kwargs):
        return self

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            d = dict(state[0] or {})
            d.update(state[1])
            state = d
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


_cache = {}


def _stub(module, name):
    key = (module, name)
    if key not in _cache:
        _cache[key] = type(name, (_Bag,), {"__module__": module, "__reduce_ex__": object.__reduce_ex__})
    return _cache[key]


class StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        top = module.split(".")[0]
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core", 1)
        if top in ("msm_we", "mdtraj", "westpa", "sklearn", "ray", "rich", "deeptime"):
            return _stub(module, name)
        return super().find_class(module, name)


def load(path):
    with open(path, "rb") as f:
        return StubUnpickler(io.BytesIO(f.read())).load()


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are committed, nothing to do")
    model = load(os.path.join(REF, "clustered.obj"))
    clusters = model.clusters
    mapper = clusters.bin_mapper
    bounds = [np.asarray(b, dtype=np.float32) for b in mapper.__dict__["_boundaries"]]
    nb = len(clusters.cluster_models)
    centers = {}
    fitted = np.zeros(nb, dtype=bool)
    for b, cm in enumerate(clusters.cluster_models):
        c = cm.__dict__.get("cluster_centers_")
        if c is not None:
            centers[f"centers_{b}"] = np.asarray(c, dtype=np.float64)
            fitted[b] = True
    remap = np.array([clusters.we_remap[b] for b in range(nb)], dtype=np.int64)
    parents, children, lens = [], [], []
    for pairs in model.pair_dtrajs:
        arr = np.array(pairs, dtype=np.int64).reshape(-1, 2)
        parents.append(arr[:, 0])
        children.append(arr[:, 1])
        lens.append(arr.shape[0])
    raw = np.load(os.path.join(REF, "fluxmatrix_raw.npy"))
    nz_i, nz_j = np.nonzero(raw)
    np.savez_compressed(
        os.path.join(OUT, "ntl9_clustered.npz"),
        boundaries=np.concatenate(bounds),
        boundary_lens=np.array([len(b) for b in bounds], dtype=np.int64),
        fitted=fitted,
        we_remap=remap,
        pair_parent=np.concatenate(parents),
        pair_child=np.concatenate(children),
        pair_lens=np.array(lens, dtype=np.int64),
        n_clusters=np.int64(model.n_clusters),
        basis_bounds=np.asarray(model._basis_pcoord_bounds, dtype=np.float64),
        target_bounds=np.asarray(model._target_pcoord_bounds, dtype=np.float64),
        flux_raw_shape=np.array(raw.shape, dtype=np.int64),
        flux_raw_nz_i=nz_i.astype(np.int64),
        flux_raw_nz_j=nz_j.astype(np.int64),
        flux_raw_sum=np.float64(raw.sum()),
        **centers,
    )
    org = load(os.path.join(REF, "organized.obj"))
    np.savez_compressed(
        os.path.join(OUT, "ntl9_downstream.npz"),
        fluxmatrix=np.load(os.path.join(REF, "fluxmatrix.npy")),
        tmatrix=np.load(os.path.join(REF, "tmatrix.npy")),
        pSS=np.load(os.path.join(REF, "pSS.npy")),
        JtargetSS=np.load(os.path.join(REF, "JtargetSS.npy")),
        indBasis=np.asarray(org.indBasis, dtype=np.int64),
        indTargets=np.asarray(org.indTargets, dtype=np.int64),
    )
    # known answers printed in the reference's unit tests
    nmm = np.array([
        [0.33380383, 0.0, 0.33455463, 0.0, 0.0, 0.33164154],
        [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        [0.33983051, 0.0, 0.32717918, 0.0, 0.0, 0.33299031],
        [0.32879530, 0.0, 0.0, 0.33194167, 0.0, 0.33926302],
        [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        [0.33247538, 0.0, 0.0, 0.33109867, 0.0, 0.33642594],
    ])
    np.savez_compressed(os.path.join(OUT, "colour_kat.npz"), nmm_tmatrix=nmm,
                        seed=np.int64(192348), n=np.int64(100000), lag=np.int64(100))
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
