"""GPU-backed ``calculate`` (reference: src/utils/calculate.py; live functions only)."""
import types

import numpy as np

from _pkg import ops as _ops


def get_believe(points2d):
    """Mean confidence of the joints with conf >= 0 (src/utils/calculate.py:8-14) -> ``pam_mean_confidence``."""
    pts = np.asarray(points2d, dtype=np.float64)
    cam = _single_camera(np.eye(3), np.zeros(3))
    return float(_ops.get_ops([cam], pts.shape[0]).mean_confidence(pts[None])[0])


def _single_camera(RK_INV, position):
    return types.SimpleNamespace(cid=0, P=np.zeros((3, 4), np.float32), RK_INV=np.asarray(RK_INV, np.float32),
                                 position=np.asarray(position, np.float64), F=np.zeros((1, 3, 3), np.float32))


def line2point_distance_3D(camera_position, directions, points3d):
    """Distance of 3-D points to rays through ``camera_position`` (src/utils/calculate.py:26-32).
    A direction d is the ray of the "pixel" (d0/d2, d1/d2) of a camera with R^-1 K^-1 = I."""
    d = np.asarray(directions, dtype=np.float64).reshape(-1, 3)
    X = np.asarray(points3d, dtype=np.float64).reshape(-1, 3)
    cam = _single_camera(np.eye(3), camera_position)
    # homogeneous pixel (u, v, 1) ~ direction: scale so the third component is 1 (|d2| > 0 assumed,
    # otherwise rotate axes); general directions are handled by permuting to the largest component
    out = np.empty(len(d))
    k = np.argmax(np.abs(d), axis=1)
    for axis in range(3):
        sel = np.nonzero(k == axis)[0]
        if len(sel) == 0:
            continue
        perm = [(axis + 1) % 3, (axis + 2) % 3, axis]
        RK = np.zeros((3, 3))
        for r, c in enumerate(perm):
            RK[c, r] = 1.0           # maps (a, b, 1) back to the original axis order
        camp = _single_camera(RK, camera_position)
        uv = np.stack([d[sel, perm[0]] / d[sel, axis], d[sel, perm[1]] / d[sel, axis]], 1)
        dist, _ = _ops.get_ops([camp], 1).ray_distance(0, uv, X[sel])
        out[sel] = dist
    return out


def line2line_distance_3D(pt1, directions1, pt2, directions2):
    """src/utils/calculate.py:20-24.  No caller anywhere in the reference (dead code); kept as plain
    numpy only so that ``from calculate import line2line_distance_3D`` (matching.py:9) still resolves."""
    n = np.cross(directions1, directions2)
    n = n / np.linalg.norm(n, axis=1).reshape(-1, 1)
    return np.abs(np.sum(n * (pt1 - pt2), axis=1))


def line_to_point_distance(a, b, c, x, y):
    """ufunc of src/utils/calculate.py:16-18: imported at matching.py:9 but never called (dead code);
    plain numpy so the import resolves."""
    return np.abs(a * x + b * y + c) / np.sqrt(np.square(a) + np.square(b))
