"""GPU-backed ``KalmanFilter`` (reference: src/tracking/KalmanFilter.py:4-65, a constant-acceleration
``cv2.KalmanFilter(9, 3)`` per 3-D joint; every use is commented out in the reference tracker).

``KalmanBank`` holds one filter per joint of a pose in device memory and steps them with one launch
(``pam_kalman9``); ``KalmanFilter`` keeps the reference's per-joint call surface on a one-filter bank."""
import numpy as np

from _pkg import ops as _ops


class KalmanBank(object):
    """``n`` filters initialised at the joints ``pts3d (n, 3)``; ``predict(pts3d=None)`` = correct (if given) + predict."""

    def __init__(self, pts3d, Hz=25, device=0):
        import torch
        p = np.asarray(pts3d, dtype=np.float64).reshape(-1, 3)
        self.Hz, self._device = Hz, int(device)
        st = np.zeros((len(p), 90), np.float32)
        st[:, :3] = p.astype(np.float32)                       # statePre = (x, y, z, 0, ...), KalmanFilter.py:46-50
        self._state = torch.from_numpy(st).to(f"cuda:{device}")

    def predict(self, pts3d=None):
        return _ops.kalman9(None if pts3d is None else np.asarray(pts3d, dtype=np.float64), self._state, self.Hz,
                            device=self._device)


class KalmanFilter(object):
    def __init__(self, pt3d, Hz=25):
        self.Hz = Hz
        self._bank = KalmanBank(np.asarray(pt3d, dtype=np.float64).reshape(1, 3), Hz)

    def predict(self, pt3d=None):
        out = self._bank.predict(None if pt3d is None else np.asarray(pt3d, dtype=np.float64).reshape(1, 3))
        return out[0].astype(np.float32)
