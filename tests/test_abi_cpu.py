"""CPU: the C-ABI shared library loads and exports every symbol include/msm_we_b200.h declares; the Python
layer refuses to run anything without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msm_we_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MWE_API\s+[\w\s\*]+?\b(mwe_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from msm_we_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 18
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
    # and the ctypes signature table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_host_only_entry_points():
    from msm_we_b200 import _lib

    assert _lib.lib.mwe_abi_version() == _lib.ABI_VERSION == 1
    # workspace-size queries are pure host functions
    assert _lib.lib.mwe_sort_workspace_bytes(1000) > 1000 * 12
    assert _lib.lib.mwe_assign_workspace_bytes(1000, 30) >= 2 * 1000 * 4
    assert _lib.lib.mwe_flux_workspace_bytes(1000) > _lib.lib.mwe_sort_workspace_bytes(1000)
    assert _lib.lib.mwe_centroid_workspace_bytes(1000, 64) > 0
    assert _lib.lib.mwe_hotpath_workspace_bytes(1000, 30) > _lib.lib.mwe_flux_workspace_bytes(1000)
    # argument validation happens before any CUDA call
    rc = _lib.lib.mwe_divide_f64(None, -1, 2.0, None)
    assert rc == -1 and b"negative" in _lib.lib.mwe_last_error()


def test_header_cites_reference_call_sites():
    text = open(HEADER).read()
    for ref in ("msm_we/stratified_clustering.py", "msm_we/_hamsm/_fluxmatrix.py", "msm_we/_hamsm/_clustering.py",
                "sklearn/cluster/_k_means_lloyd.pyx", "sklearn/cluster/_k_means_minibatch.pyx", "msm_we/nmm.py",
                "msm_we/msm_we.py"):
        assert ref in text


def test_no_cpu_fallback():
    import torch
    from msm_we_b200 import ops
    from msm_we_b200.engine import require_cuda

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        require_cuda()
    with pytest.raises(TypeError):
        ops.centers_sqnorm(torch.zeros(4, 3, dtype=torch.float64))
    with pytest.raises(TypeError):
        ops.flux_accumulate(torch.zeros(3, dtype=torch.int64), torch.zeros(3, dtype=torch.int64), None, 5)
    from msm_we_b200.stratified_clustering import BinClusterModel

    m = BinClusterModel(n_clusters=2)
    with pytest.raises(RuntimeError):
        m.partial_fit(np.zeros((4, 2)))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "msm_we_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f
