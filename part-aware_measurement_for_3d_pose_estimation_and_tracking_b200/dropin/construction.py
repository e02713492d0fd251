"""GPU-backed ``construction`` (reference: src/utils/construction.py:9-31, 64-131)."""
from math import exp

import numpy as np

from _pkg import ops as _ops


def SVD_pose_kernel_jf(cameras, Ts, pose_mat, lambda_t, remains, joints_views, next_pose=None):
    """Per-joint view-filtered weighted DLT (src/utils/construction.py:89-114) -> (J, 3)."""
    pose = np.asarray(pose_mat, dtype=np.float64)
    Vt, J = pose.shape[0], pose.shape[1]
    keep = (np.asarray(remains)[:, ::2] == 1).astype(np.uint8)
    w = np.array([exp(-lambda_t * T) for T in Ts])
    o = _ops.get_ops(list(cameras), J)
    out = o.triangulate(np.arange(Vt), pose, w, keep=keep[None],
                        next_pose=None if next_pose is None else np.asarray(next_pose, dtype=np.float64)[None])
    listed = np.zeros(J, bool)
    for js in joints_views:
        listed[list(js)] = True
    out[~listed] = 0.0
    return out


def SVD_pose_kernel_parallel(cameras, Ts, pose_mat, lambda_t):
    """All views, all joints (src/utils/construction.py:116-131) -> (J, 3)."""
    pose = np.asarray(pose_mat, dtype=np.float64)
    w = np.array([exp(-lambda_t * T) for T in Ts])
    return _ops.get_ops(list(cameras), pose.shape[1]).triangulate(np.arange(pose.shape[0]), pose, w)


def SVD_pose_kernel(cameras, Ts, joints, remains, lambda_t, next_pose=None):
    """Older per-joint form (src/utils/construction.py:64-87): ``joints[j][view] = (v, u, ...)``,
    ``remains[j]`` = list of view indices -> list of (3,) arrays."""
    J, Vt = len(joints), len(cameras)
    pose = np.zeros((Vt, J, 3))
    keep = np.zeros((J, Vt), np.uint8)
    for j, (joint, remain) in enumerate(zip(joints, remains)):
        for v in remain:
            pose[v, j, :2] = np.asarray(joint[v], dtype=np.float64)[:2]
            keep[j, v] = 1
    w = np.array([exp(-lambda_t * T) for T in Ts])
    nx = None if next_pose is None else np.asarray(next_pose, dtype=np.float64)[None]
    out = _ops.get_ops(list(cameras), J).triangulate(np.arange(Vt), pose, w, keep=keep[None], next_pose=nx)
    res = []
    for j in range(J):
        if keep[j].sum() <= 1 and next_pose is None:
            res.append(np.array([None, None, None]))
        else:
            res.append(out[j])
    return res


def top_down_pose_kernel(cameras, poses2d, weight2d=None):
    """Pair-wise triangulation, best pair by summed reprojection error (src/utils/construction.py:9-31)
    -> (pose3d (J, 3), weight of the winning pair)."""
    p = np.asarray([np.asarray(q, dtype=np.float64) for q in poses2d])
    pose3d, pair = _ops.get_ops(list(cameras), p.shape[1]).top_down(np.arange(p.shape[0]), p)
    return pose3d, (weight2d[int(pair[0])] + weight2d[int(pair[1])]) / 2
