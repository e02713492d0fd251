"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's alternative smoothers and pair-wise triangulation
(kept in the reference's API, every call commented out or without a caller).

  one_euro_*            src/tracking/OneEuroFilter.py:12-77   (LowPassFilter + OneEuroFilter, float64 scalars)
  top_down_pose_kernel  src/utils/construction.py:9-31        (cv2.triangulatePoints over all camera pairs, the pair
                                                               with the smallest summed reprojection error wins)

Checked bit for bit against the unmodified reference modules (tests/test_alt_ops.py)."""
from __future__ import annotations

import math

import numpy as np


class OneEuroState:
    """State of ONE filter (the reference keeps it in two LowPassFilter objects, OneEuroFilter.py:51-53)."""

    def __init__(self, freq, mincutoff=1.0, beta=0.0, dcutoff=1.0):
        if freq <= 0 or mincutoff <= 0 or dcutoff <= 0:
            raise ValueError("freq, mincutoff and dcutoff should be >0")          # OneEuroFilter.py:42-47
        self.freq, self.mincutoff, self.beta, self.dcutoff = float(freq), float(mincutoff), float(beta), float(dcutoff)
        self.x_prev = None      # LowPassFilter.__y of the value filter
        self.s_x = None         # LowPassFilter.__s of the value filter
        self.s_dx = None        # ... of the derivative filter
        self.lasttime = None


def one_euro_alpha(freq, cutoff):
    te = 1.0 / freq                                   # OneEuroFilter.py:55-58
    tau = 1.0 / (2 * math.pi * cutoff)
    return 1.0 / (1.0 + tau / te)


def one_euro_step(st: OneEuroState, x, timestamp=None):
    """OneEuroFilter.__call__ (OneEuroFilter.py:60-77)."""
    if x is None:
        return x
    if st.lasttime and timestamp:
        st.freq = 1.0 / (timestamp - st.lasttime)
    st.lasttime = timestamp
    dx = 0.0 if st.x_prev is None else (x - st.x_prev) * st.freq
    a_d = one_euro_alpha(st.freq, st.dcutoff)
    edx = dx if st.s_dx is None else a_d * dx + (1.0 - a_d) * st.s_dx             # LowPassFilter.__call__ :24-33
    st.s_dx = edx
    cutoff = st.mincutoff + st.beta * math.fabs(edx)
    a = one_euro_alpha(st.freq, cutoff)
    s = x if st.s_x is None else a * x + (1.0 - a) * st.s_x
    st.x_prev, st.s_x = x, s
    return s


def top_down_pose_kernel(cameras, poses2d, weight2d=None):
    """construction.py:9-31.  cameras: objects with `.P (3,4)`; poses2d: per camera (J, 2) pixel coordinates in the
    order cv2.triangulatePoints expects; weight2d: per camera weight.  Returns (pose3d (J,3), weight of the pair)."""
    import cv2
    poses3d, weight3d, reproj_error = [], [], []
    for i in range(len(poses2d)):
        for j in range(i + 1, len(poses2d)):
            homo = cv2.triangulatePoints(cameras[i].P, cameras[j].P, poses2d[i].T, poses2d[j].T)
            poses3d.append(homo[:3] / homo[3])
            weight3d.append((weight2d[i] + weight2d[j]) / 2)
            err = 0
            for camera, pk in zip(cameras, poses2d):
                ph = camera.P @ homo
                pr = ph[:2] / (ph[2] + 10e-6)
                err += np.linalg.norm(pr.T - pk)
            reproj_error.append(err)
    idx = int(np.argmin(reproj_error))
    return poses3d[idx].T, weight3d[idx]
