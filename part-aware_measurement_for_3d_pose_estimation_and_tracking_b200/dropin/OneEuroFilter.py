"""GPU-backed ``OneEuroFilter`` (reference: src/tracking/OneEuroFilter.py:12-77).

The reference filters one python float per object and IterTrack keeps 3 x J of them per track
(src/tracking/IterativeTracker.py:231-237; every call is commented out there).  ``OneEuroBank`` is that set as ONE
device-resident bank -- all channels share the time stamps, one kernel launch per step (``pam_one_euro``) -- and
``OneEuroFilter`` keeps the reference's scalar call surface on top of a one-channel bank (same arithmetic in the same
order, bit-identical results; for real use prefer the bank)."""
import numpy as np

from _pkg import ops as _ops


class OneEuroBank(object):
    """``n`` One-Euro filters with common time stamps; ``bank(x, timestamp)`` filters the (n,) array ``x``."""

    def __init__(self, n, freq, mincutoff=1.0, beta=0.0, dcutoff=1.0, device=0):
        if freq <= 0:
            raise ValueError("freq should be >0")
        if mincutoff <= 0:
            raise ValueError("mincutoff should be >0")
        if dcutoff <= 0:
            raise ValueError("dcutoff should be >0")
        import torch
        self._freq, self._mincutoff, self._beta, self._dcutoff = float(freq), float(mincutoff), float(beta), float(dcutoff)
        self._lasttime = None
        self._device = int(device)
        self._state = torch.zeros((int(n), 4), dtype=torch.float64, device=f"cuda:{device}")

    def __call__(self, x, timestamp=None):
        if x is None:
            return x
        if self._lasttime and timestamp:                       # OneEuroFilter.py:64-65
            self._freq = 1.0 / (timestamp - self._lasttime)
        self._lasttime = timestamp
        return _ops.one_euro(x, self._state, self._freq, self._mincutoff, self._beta, self._dcutoff, self._device)


class OneEuroFilter(object):
    def __init__(self, freq, mincutoff=1.0, beta=0.0, dcutoff=1.0):
        self._bank = OneEuroBank(1, freq, mincutoff, beta, dcutoff)

    def __call__(self, x, timestamp=None):
        if x is None:
            return x
        return float(self._bank(np.array([x], dtype=np.float64), timestamp)[0])
