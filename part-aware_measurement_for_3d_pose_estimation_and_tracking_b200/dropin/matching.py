"""GPU-backed ``matching`` (reference: src/utils/matching.py; live functions only)."""
import numpy as np

from _pkg import ops as _ops
from calculate import _single_camera, get_believe, line2point_distance_3D, line2line_distance_3D, line_to_point_distance  # noqa: F401


def back_project_ray(RK_INV, camera_position, points):
    """Unit rays of (u, v) pixels (src/utils/matching.py:10-17)."""
    pts = np.asarray(points, dtype=np.float64)
    cam = _single_camera(RK_INV, camera_position)
    _, dirs = _ops.get_ops([cam], 1).ray_distance(0, pts[:, :2], None, want_dirs=True)
    return dirs


def epipolar_distance(cam1, person1, cam2, person2):
    """(J, 2) ``[d(x1, F x2), d(x2, F^T x1)]`` (src/utils/matching.py:50-91)."""
    p1, p2 = np.asarray(person1, dtype=np.float64), np.asarray(person2, dtype=np.float64)
    if len(p1) == 0:
        return []
    o = _ops.get_ops([cam1, cam2], len(p1))
    return o.epipolar_distance([0], p1[None], [1], p2[None])[0]


def epipolar_affinity(cameras, sub_imgid2cam, pose_mat, num_joints):
    """-> ((M, M) float32 mean distance, (M, M, J) float32) (src/utils/matching.py:93-113)."""
    pose = np.array(pose_mat, dtype=np.float64)
    o = _ops.get_ops(list(cameras), num_joints)
    return o.epipolar_allpairs(np.asarray(sub_imgid2cam), pose, want_dist=True)


def epipolar_affinity_parallel(cameras, sub_imgid2cam, pose_mat, num_joints):
    """-> ((M, M) float64 mean distance, (M, M, J) float64) (src/utils/matching.py:115-151)."""
    pose = np.array(pose_mat, dtype=np.float64)
    o = _ops.get_ops(list(cameras), num_joints)
    return o.epipolar_pairs(np.asarray(sub_imgid2cam), pose)


def Greedy_matching(cameras, pose_mat=None, affinity_mat=None, costs=None, next_pose=None, mode='update'):
    """-> (matched_list, binary_list, affinity_mat) (src/utils/matching.py:243-295)."""
    n = affinity_mat.shape[0]
    o = _ops.get_ops(list(cameras), 1)
    if mode == 'update':
        uv = np.flip(np.asarray(pose_mat, dtype=np.float64).reshape(n, -1)[:, :2], axis=1)
        keep = o.view_filter(np.arange(n), np.asarray(affinity_mat)[None], 'update', uv[None],
                             np.asarray(next_pose, dtype=np.float64).reshape(1, 3))[0]
    else:
        keep = o.view_filter(np.arange(n), np.asarray(affinity_mat)[None], 'init')[0]
    matched_list = np.nonzero(keep)[0]
    binary_list = np.repeat(keep.astype(int), 2)
    return matched_list, binary_list, affinity_mat


def BIP_matching(model, cameras, dimGroup, pose_mat=None, num_joints=17, threshold=40):
    """All-detections clustering front end (src/utils/matching.py:234-241): camera index per detection from the
    ``dimGroup`` offsets, all-pairs epipolar affinity on the device (``pam_epipolar_allpairs``, float32 like the
    reference), ``1 - affinity / threshold``, and the CALLER's solver object decides the clusters
    (``model.solve``; the reference instantiates ``GLPKSolver`` from tracking/binary_integer_programming.py, which needs
    cvxopt and a pre-1.11 scipy simplex -- neither is part of this package).  -> (matched_list, sub_imgid2cam)."""
    sub_imgid2cam = np.zeros(dimGroup[-1] if dimGroup[-1] - 1 >= 0 else 0, dtype=np.int32)
    for idx, i in enumerate(range(len(dimGroup) - 1)):
        sub_imgid2cam[dimGroup[i]:dimGroup[i + 1]] = idx
    affinity_mat, _ = epipolar_affinity(cameras, sub_imgid2cam, pose_mat, num_joints)
    affinity_mat = 1 - affinity_mat / threshold
    matched_list = model.solve(affinity_mat.astype(np.double))
    return matched_list, sub_imgid2cam
