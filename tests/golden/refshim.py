"""Loads the UNMODIFIED reference package from /root/reference inside this container, so its own code can be
executed to produce golden fixtures (tests/golden/make_reference_fixtures.py).

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (not on the GPU box): the fixtures it
produces are committed, this loader is not needed to run the tests.

``import msm_we`` needs third-party packages that are not installed here (mdtraj, ray, westpa, h5py, deeptime,
matplotlib) and one scipy symbol that newer scipy dropped.  None of them holds arithmetic of the hot path, so
they are replaced by minimal stand-ins registered in ``sys.modules`` BEFORE the import; every line of
``/root/reference/msm_we`` then runs as written:

* ``ray``      -- ``remote`` / ``put`` / ``get`` / ``wait`` executed synchronously in submission order (the reference
                  sums flux matrices in completion order; synchronous execution = iteration order).
* ``h5py``     -- ``File(name)`` serving numpy arrays registered under that name in ``FILES`` with the WESTPA layout
                  (``iterations/iter_%08d/{seg_index,pcoord,auxdata/<auxpath>}``; msm_we/_hamsm/_data.py:807-932,
                  westpa_plugins/augmentation_driver.py:173-180).
* ``westpa``   -- ``core.binning.RectilinearBinMapper`` / ``VoronoiBinMapper`` restated from westpa's published
                  semantics (float32 coordinates, ``lower <= x < upper``, last dimension fastest, ValueError when out
                  of range; UNVERIFIED against westpa source, which is not in this image), ``analysis.Run`` returning
                  the bin mapper registered for a file, ``tools.binning`` empty.
* ``mdtraj``, ``deeptime``, ``matplotlib`` -- empty shells (never reached on this path: topology is passed as a dict,
                  ``dim_reduce_method="none"``, no plotting).
* ``scipy.sparse.sputils.isdense`` -- re-created (msm_we/utils.py:16 imports it; scipy >= 1.8 moved it).
* numpy >= 1.24 refuses ragged ``np.array(list_of_arrays)``; ``get_cluster_centers`` (_clustering.py:1594) relies on
  the older behaviour (object array).  ``ragged_numpy()`` patches the ``np`` name of that one module with a proxy whose
  ``array`` retries with ``dtype=object``.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"

FILES = {}        # name -> {dataset path without leading '/': ndarray}
BIN_MAPPERS = {}  # name -> bin mapper returned by westpa.analysis.Run(name).iteration(n).bin_mapper


# ---------------------------------------------------------------------------------------------- h5py
class _Dataset(np.ndarray):
    """ndarray with the h5py.Dataset idioms the reference uses (``dset[:]``, ``dset["weight"]``) and
    ``Dataset.read_direct`` with h5py's contract: the destination must be a C-contiguous, writable array and the two
    selections (``numpy.s_`` outputs) must have the same shape."""

    def read_direct(self, dest, source_sel=None, dest_sel=None):
        if not (isinstance(dest, np.ndarray) and dest.flags.c_contiguous and dest.flags.writeable):
            raise TypeError("Destination array must be C-contiguous and writable")
        src = np.asarray(self)[source_sel if source_sel is not None else np.s_[...]]
        view = dest[dest_sel if dest_sel is not None else np.s_[...]]
        if view.shape != src.shape:
            raise TypeError(f"Can't broadcast {src.shape} -> {view.shape}")
        view[...] = src


class FakeH5File:
    def __init__(self, name, mode="r", **kwargs):
        if name not in FILES:
            raise OSError(f"Unable to open file (no fake HDF5 file registered as {name!r})")
        self._d = FILES[name]
        self.filename = name

    @staticmethod
    def _key(k):
        return k.strip("/")

    def __contains__(self, k):
        k = self._key(k)
        return k in self._d or any(p.startswith(k + "/") for p in self._d)

    def __getitem__(self, k):
        k = self._key(k)
        if k in self._d:
            return self._d[k].view(_Dataset)
        sub = {p[len(k) + 1:]: v for p, v in self._d.items() if p.startswith(k + "/")}
        if not sub:
            raise KeyError(f"Unable to open object (object '{k}' doesn't exist)")
        g = FakeH5File.__new__(FakeH5File)
        g._d, g.filename = sub, self.filename
        return g

    def keys(self):
        return sorted({p.split("/")[0] for p in self._d})

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


SEG_INDEX_DTYPE = np.dtype([("weight", np.float64), ("parent_id", np.int64), ("wtg_n_parents", np.uint32),
                            ("wtg_offset", np.uint32), ("cputime", np.float64), ("walltime", np.float64),
                            ("endpoint_type", np.uint8), ("status", np.uint8)])


def register_we_file(name, iterations, auxpath="coord", bin_mapper=None):
    """``iterations``: list over WE iterations 1..n of dicts with ``weights [S]``, ``pcoord [S, pcoord_len, P]``,
    ``coords [S, 2, nAtoms, coord_ndim]`` and optionally ``parent_id [S]``.  As in a real west.h5, one more (empty
    of dynamics) iteration group is appended: the reference treats an iteration as complete only when the next one
    exists (_data.py:866-869)."""
    d = {}
    for i, it in enumerate(iterations, start=1):
        S = len(it["weights"])
        seg = np.zeros(S, dtype=SEG_INDEX_DTYPE)
        seg["weight"] = it["weights"]
        seg["parent_id"] = it.get("parent_id", np.arange(S))
        seg["endpoint_type"] = 1
        seg["status"] = 2
        g = f"iterations/iter_{i:08d}"
        d[f"{g}/seg_index"] = seg
        d[f"{g}/pcoord"] = np.ascontiguousarray(it["pcoord"], dtype=np.float64)
        d[f"{g}/auxdata/{auxpath}"] = np.ascontiguousarray(it["coords"], dtype=np.float64)
    last = len(iterations) + 1
    d[f"iterations/iter_{last:08d}/seg_index"] = np.zeros(1, dtype=SEG_INDEX_DTYPE)
    d[f"iterations/iter_{last:08d}/pcoord"] = np.zeros((1,) + tuple(np.shape(iterations[-1]["pcoord"])[1:]))
    FILES[name] = d
    if bin_mapper is not None:
        BIN_MAPPERS[name] = bin_mapper


def fake_h5py_module():
    m = types.ModuleType("h5py")
    m.File = FakeH5File
    m.__fake__ = True
    return m


# ---------------------------------------------------------------------------------------------- westpa
class RectilinearBinMapper:
    """westpa.core.binning.RectilinearBinMapper, restated (see module docstring)."""

    def __init__(self, boundaries):
        self.boundaries = [np.asarray(b, dtype=np.float32) for b in boundaries]
        self.ndim = len(self.boundaries)
        self.nbins = int(np.prod([len(b) - 1 for b in self.boundaries]))
        self.labels = [str(i) for i in range(self.nbins)]

    def assign(self, coords, mask=None, output=None):
        coords = np.asarray(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords[:, None]
        c32 = coords.astype(np.float32)
        index = np.zeros(coords.shape[0], dtype=np.uint16)
        for d, b in enumerate(self.boundaries):
            pos = np.searchsorted(b, c32[:, d], side="right") - 1
            bad = (pos < 0) | (pos >= len(b) - 1) | np.isnan(c32[:, d])
            if bad.any():
                raise ValueError("coordinate outside of bin space")
            index = (index.astype(np.int64) * (len(b) - 1) + pos).astype(np.uint16)
        return index


class VoronoiBinMapper:
    """westpa.core.binning.VoronoiBinMapper: ``dfunc(coord, centers)`` -> distances, nearest centre wins."""

    def __init__(self, dfunc, centers, dfargs=None, dfkwargs=None):
        self.dfunc = dfunc
        self.centers = np.asarray(centers)
        self.nbins = self.centers.shape[0]
        self.ndim = self.centers.shape[1]
        self.dfargs = dfargs or ()
        self.dfkwargs = dfkwargs or {}
        self.labels = [str(i) for i in range(self.nbins)]

    def assign(self, coords, mask=None, output=None):
        coords = np.asarray(coords)
        out = np.empty(len(coords), dtype=np.uint16)
        for i, c in enumerate(coords):
            out[i] = np.argmin(self.dfunc(c, self.centers, *self.dfargs, **self.dfkwargs))
        return out


class _Run:
    def __init__(self, name):
        self.name = name

    def iteration(self, n):
        return types.SimpleNamespace(bin_mapper=BIN_MAPPERS[self.name])


# ---------------------------------------------------------------------------------------------- ray
class _RemoteFunction:
    def __init__(self, fn):
        self._fn = fn
        self.__name__ = getattr(fn, "__name__", "remote")

    def __get__(self, obj, objtype=None):   # the reference decorates methods and calls self.f.remote(...)
        return self

    def remote(self, *a, **k):
        return self._fn(*a, **k)

    def __call__(self, *a, **k):
        return self._fn(*a, **k)


def _ray_module():
    ray = types.ModuleType("ray")

    def remote(*args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return _RemoteFunction(args[0])
        return lambda fn: _RemoteFunction(fn)

    ray.remote = remote
    ray.put = lambda x: x
    ray.get = lambda x: x
    ray.wait = lambda ids, num_returns=1, timeout=None: (list(ids[:num_returns]), list(ids[num_returns:]))
    ray.is_initialized = lambda: True
    ray.init = lambda *a, **k: None
    ray.shutdown = lambda *a, **k: None
    ray.available_resources = lambda: {"CPU": float(os.cpu_count() or 1)}
    ray.cluster_resources = ray.available_resources
    return ray


class _Shell:
    def __init__(self, *a, **k):
        pass


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_LOADED = None


def load_reference():
    """Returns the reference's ``msm_we`` package (imported from /root/reference with the stand-ins above)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "msm_we")):
        raise RuntimeError(f"{REFERENCE_ROOT}/msm_we not found: reference fixtures can only be generated in the build container")
    if "msm_we" in sys.modules:
        raise RuntimeError("a module named msm_we is already imported")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import scipy.sparse

    sys.modules["ray"] = _ray_module()
    sys.modules["h5py"] = fake_h5py_module()
    md = _stub("mdtraj", Trajectory=type("Trajectory", (_Shell,), {}), Topology=type("Topology", (_Shell,), {}))
    md.load = md.load_prmtop = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("mdtraj is a stand-in here"))
    westpa = _stub("westpa")
    westpa.rc = types.SimpleNamespace(pstatus=lambda *a, **k: None)
    westpa.analysis = _stub("westpa.analysis", Run=_Run)
    westpa.core = _stub("westpa.core")
    westpa.core.binning = _stub("westpa.core.binning", RectilinearBinMapper=RectilinearBinMapper,
                                VoronoiBinMapper=VoronoiBinMapper)
    westpa.core.extloader = _stub("westpa.core.extloader")
    westpa.tools = _stub("westpa.tools")
    westpa.tools.binning = _stub("westpa.tools.binning")
    dt = _stub("deeptime")
    dt.decomposition = _stub("deeptime.decomposition", TICA=_Shell, VAMP=_Shell)
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    try:
        from scipy.sparse.sputils import isdense  # noqa: F401
    except ImportError:
        _stub("scipy.sparse.sputils", isdense=lambda x: isinstance(x, np.ndarray))
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import msm_we
    finally:
        sys.path.remove(REFERENCE_ROOT)
    _LOADED = msm_we
    return msm_we


class _RaggedNumpy:
    """``np`` look-alike whose ``array`` builds an object array where numpy >= 1.24 raises for ragged input."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def array(obj, *a, **k):
        try:
            return np.array(obj, *a, **k)
        except ValueError:
            out = np.empty(len(obj), dtype=object)
            for i, o in enumerate(obj):
                out[i] = o
            return out


def ragged_numpy(module):
    """Patch ``module.np`` (see module docstring)."""
    module.np = _RaggedNumpy()
