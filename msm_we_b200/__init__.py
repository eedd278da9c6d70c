"""msm_we_b200 -- B200-native (sm_100a) implementation of msm_we's discretization + flux hot path.

Importing the package loads ``libmsm_we_b200.so`` through ctypes; if the CUDA extension has not
been built the import fails (there is no CPU fallback anywhere in this package).
"""
from . import _lib  # noqa: F401  (fails loudly when the extension is missing)

__version__ = "0.1.0"
