"""Page-locking of host arrays a model already holds, so chunks can go to the device without a staging copy.

The reference hands coordinates around as numpy arrays (msm_we/_hamsm/_data.py:557-618).  Copying them into a
pinned staging buffer first costs more than the PCIe transfer itself (one core moves ~8 GB/s, eight ~30 GB/s, the
link 48 GB/s), and the same arrays are read again by every later pass over the data (clustering, discretization,
block validation).  ``HostPins.ensure(arr)`` page-locks the buffer that owns ``arr`` once (``mwe_host_register``)
and keeps it locked until the owner is garbage-collected; afterwards ``torch.from_numpy(arr)`` copies
asynchronously at link speed.  Anything that cannot be locked (not C-contiguous fp64, over the budget,
overlapping an existing registration, ``MSM_WE_B200_PIN_HOST=0``) simply reports False and the caller stages it.
"""
from __future__ import annotations

import os
import weakref

import numpy as np


def _default_budget():
    try:
        return int(os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") // 4)
    except (ValueError, OSError, AttributeError):
        return 8 << 30


class HostPins:
    def __init__(self):
        self._owners = {}            # id(owner array) -> (ptr, nbytes)
        self._failed = set()         # id(owner) that could not be registered (do not retry every pass)
        self.bytes = 0
        self.enabled = os.environ.get("MSM_WE_B200_PIN_HOST", "1") != "0"
        self.budget = int(os.environ.get("MSM_WE_B200_PIN_HOST_BYTES", "0")) or _default_budget()

    @staticmethod
    def _owner(arr):
        b = arr
        while isinstance(b.base, np.ndarray):
            b = b.base
        return b

    def _release(self, key, ptr, nbytes):
        from . import _lib

        if self._owners.pop(key, None) is not None:
            self.bytes -= nbytes
            try:
                _lib.lib.mwe_host_unregister(ptr)
            except Exception:       # interpreter shutdown
                pass

    def ensure(self, arr) -> bool:
        """True when ``arr`` (C-contiguous float64) lies in page-locked memory after the call."""
        if not self.enabled or not isinstance(arr, np.ndarray) or arr.dtype != np.float64 or not arr.flags.c_contiguous \
                or arr.nbytes == 0:
            return False
        owner = self._owner(arr)
        key = id(owner)
        if key in self._owners:
            return True
        if key in self._failed or not owner.flags.owndata or not (owner.flags.c_contiguous or owner.flags.f_contiguous):
            return False
        from . import _lib

        nbytes = owner.nbytes
        if self.bytes + nbytes > self.budget or _lib.lib.mwe_host_register(owner.ctypes.data, nbytes) != 0:
            self._failed.add(key)
            weakref.finalize(owner, self._failed.discard, key).atexit = False
            return False
        self._owners[key] = (owner.ctypes.data, nbytes)
        self.bytes += nbytes
        # numpy clears weak references at the start of deallocation, while the buffer still exists
        weakref.finalize(owner, self._release, key, owner.ctypes.data, nbytes).atexit = False
        return True


PINS = HostPins()
