"""TEST INFRASTRUCTURE ONLY -- J-parametrised CPU restatement of the reference hot path.

This file is the *oracle* the CUDA path is checked against.  It restates, in numpy, the live
call graph of the reference's per-frame geometric path and
generalises ONLY the five sites where the reference hard-wires 17 joints:

  ========================================  =====================================================
  here                                       reference (path:line under /root/reference)
  ========================================  =====================================================
  ``Camera`` / ``Camera.project_tracks``     src/ivclabpose.py:35-46, 91-98   (``reshape(-1, 17, 2)``)
  ``build_cameras``                          src/ivclabpose.py:162-181
  ``mean_confidence``                        src/utils/calculate.py:8-14
  ``ray_point_distance``                     src/utils/calculate.py:26-32
  ``pixel_rays``                             src/utils/matching.py:10-17
  ``epilines`` (cv2 restated)                OpenCV ``cv::computeCorrespondEpilines`` (fundam.cpp),
                                             call sites src/utils/matching.py:69-73
  ``epipolar_distance``                      src/utils/matching.py:50-91
  ``epipolar_affinity``                      src/utils/matching.py:93-113
  ``epipolar_affinity_parallel``             src/utils/matching.py:115-151
  ``greedy_view_filter``                     src/utils/matching.py:243-295
  ``dlt_joint_filtered``                     src/utils/construction.py:89-114 (``np.zeros((17, 3))``)
  ``dlt_all_views``                          src/utils/construction.py:116-131
  ``dlt_per_joint``                          src/utils/construction.py:64-87
  ``Hypothesis``                             src/tracking/hypothesis.py:9-77
  ``Track``                                  src/tracking/IterativeTracker.py:182-395
                                             (``not_arm`` / ``[9,10]`` at :380-382)
  ``Tracker``                                src/tracking/IterativeTracker.py:34-180 (``> 10`` at :145)
  ========================================  =====================================================

Floating-point contract: every array operation below is issued with the same numpy / scipy
primitive, operand dtypes and operand shapes as the reference line it cites, so that the results
are bit-identical to the unmodified reference at J = 17 -- ``tests/test_oracle_vs_reference.py``
proves this in the build container and ``tests/golden/`` pins it for the GPU box.

Parity pinning: the reference ships no tests, fixtures or golden vectors,
so the oracle is pinned against *outputs of the reference itself run in the build container*
(``tests/golden/make_golden.py``); see DESIGN.md "Oracle".

Third-party arithmetic on the path (not vendored by the reference; this container's versions):
``numpy.linalg.svd`` (LAPACK gesdd), ``scipy.optimize.linear_sum_assignment``,
``scipy.ndimage.gaussian_filter1d``, ``cv2.computeCorrespondEpilines`` (restated in ``epilines``),
``torch.inverse`` (camera set-up).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product package never does.
"""
from __future__ import annotations

from math import exp
from typing import List, Optional, Sequence

import numpy as np
import numpy.linalg as la
from scipy.ndimage import gaussian_filter1d
from scipy.optimize import linear_sum_assignment

TENTATIVE, CONFIRMED, DELETED = 1, 2, 3


# ----------------------------------------------------------------------------------------------
# cameras
# ----------------------------------------------------------------------------------------------


class Camera:
    """Per-camera constants (src/ivclabpose.py:35-46).  ``P, K, RT, F, RK_INV`` are float32,
    ``position`` is float64 -- exactly the dtypes the reference ends up with."""

    def __init__(self, cid, P, K, RT, F, w=640, h=480):
        self.cid = cid
        self.P, self.K, self.RT, self.F = P, K, RT, F
        self.w, self.h = w, h
        self.RK_INV = la.inv(RT[:, :3]) @ la.inv(K)
        self.position = la.inv(np.vstack([RT, [0, 0, 0, 1]]))[:3, 3]

    def project_tracks(self, points3d):
        """(n, J, 3) world joints -> (n, J, 2) pixels as (v, u) (src/ivclabpose.py:91-98)."""
        n, J = points3d.shape[0], points3d.shape[1]
        homo = np.concatenate([points3d, np.ones((n, J, 1))], axis=2)
        pr = np.transpose(self.P @ homo.reshape(-1, 4).T)
        uv = pr[:, :2] / pr[:, 2].reshape(-1, 1)
        return np.flip(uv, axis=1).reshape(-1, J, 2)

    # reference spelling, so reference-style call sites keep working on oracle cameras
    projectPoints_parallel = project_tracks


def build_cameras(P, K, RT, w=640, h=480) -> List[Camera]:
    """``GetCameraParameters`` (src/ivclabpose.py:162-181): float32 casts and the V x V tensor of
    fundamental matrices, evaluated with float32 torch CPU ops like the reference does."""
    import torch

    P = np.asarray(P).astype(np.float32)
    K = np.asarray(K).astype(np.float32)
    RT = np.asarray(RT).astype(np.float32)
    V = len(P)

    def cross_matrix(x):
        return torch.tensor([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]])

    def fundamental(K0, RT0, K1, RT1):
        R0, T0, R1, T1 = RT0[:, :3], RT0[:, 3], RT1[:, :3], RT1[:, 3]
        return torch.inverse(K0).t() @ (R0 @ R1.t()) @ K1.t() @ cross_matrix(
            K1 @ R1 @ R0.t() @ (T0 - R0 @ R1.t() @ T1))

    F = torch.zeros(V, V, 3, 3)
    for a in range(V):
        for b in range(V):
            F[a, b] += fundamental(torch.tensor(K[a]), torch.tensor(RT[a]), torch.tensor(K[b]), torch.tensor(RT[b]))
            if F[a, b].sum() == 0:
                F[a, b] += 1e-12
    F = F.numpy()
    return [Camera(j, P[j], K[j], RT[j], F[j], w=w, h=h) for j in range(V)]


# ----------------------------------------------------------------------------------------------
# small helpers (src/utils/calculate.py)
# ----------------------------------------------------------------------------------------------


def mean_confidence(points2d):
    """``get_believe`` (src/utils/calculate.py:8-14)."""
    kept = [p[2] for p in points2d if p[2] >= 0]
    return np.mean(kept)


def ray_point_distance(camera_position, directions, points3d):
    """``line2point_distance_3D`` (src/utils/calculate.py:26-32)."""
    x0 = points3d.astype(float)
    x1 = camera_position
    x2 = camera_position + directions
    cr = np.cross(x2 - x1, x1 - x0)
    return la.norm(cr, axis=1) / la.norm(x2 - x1, axis=1)


def line_line_distance(pt1, directions1, pt2, directions2):
    """``line2line_distance_3D`` (src/utils/calculate.py:20-24)."""
    n = np.cross(directions1, directions2)
    n = n / la.norm(n, axis=1).reshape(-1, 1)
    return np.abs(np.sum(n * (pt1 - pt2), axis=1))


def pixel_rays(RK_INV, camera_position, points):
    """``back_project_ray`` (src/utils/matching.py:10-17): unit directions of (u, v) pixels."""
    n = len(points)
    homo = np.concatenate([points[:, :2], np.ones((n, 1))], axis=1)
    d = (np.repeat([RK_INV], n, axis=0) @ homo.reshape(n, 3, 1))[:, :3].reshape(n, 3)
    return d / la.norm(d, axis=1).reshape(n, 1)


# ----------------------------------------------------------------------------------------------
# epipolar geometry (src/utils/matching.py)
# ----------------------------------------------------------------------------------------------


def epilines(points, which_image, F):
    """Restatement of ``cv2.computeCorrespondEpilines`` for float64 points.

    OpenCV (modules/calib3d/src/fundam.cpp) converts ``F`` to double, transposes it when
    ``which_image == 2`` and evaluates, per point, ``a = f0*x + f1*y + f2`` (likewise b, c),
    ``nu = a*a + b*b``, ``nu = nu ? 1/sqrt(nu) : 1`` and returns ``(a*nu, b*nu, c*nu)``.
    ``tests/test_oracle_vs_reference.py`` checks this bit-for-bit against cv2."""
    f = np.asarray(F, dtype=np.float64)
    if which_image == 2:
        f = f.T
    x, y = points[:, 0], points[:, 1]
    a = f[0, 0] * x + f[0, 1] * y + f[0, 2]
    b = f[1, 0] * x + f[1, 1] * y + f[1, 2]
    c = f[2, 0] * x + f[2, 1] * y + f[2, 2]
    nu = a * a + b * b
    with np.errstate(divide="ignore"):
        nu = np.where(nu != 0, 1.0 / np.sqrt(nu), 1.0)
    return np.stack([a * nu, b * nu, c * nu], axis=1)


def epipolar_distance(cam1, person1, cam2, person2):
    """(J, 2) ``[d(x1, F x2), d(x2, F^T x1)]`` in pixels (src/utils/matching.py:50-91)."""
    J = len(person1)
    F = cam1.F[cam2.cid]
    p1 = np.array(np.flip(person1[:, :2], axis=1))
    p2 = np.array(np.flip(person2[:, :2], axis=1))
    if len(p1) == 0:
        return []
    l_in_2 = epilines(p1, 2, F)
    l_in_1 = epilines(p2, 1, F)
    h1 = np.concatenate([p1, np.ones((J, 1))], axis=1)
    h2 = np.concatenate([p2, np.ones((J, 1))], axis=1)
    d1 = np.abs(np.sum(h1 * l_in_1, axis=1)) / np.sqrt(np.sum(l_in_1[:, :2] ** 2, axis=1))
    d2 = np.abs(np.sum(h2 * l_in_2, axis=1)) / np.sqrt(np.sum(l_in_2[:, :2] ** 2, axis=1))
    return np.hstack([d1.reshape(-1, 1), d2.reshape(-1, 1)])


def epipolar_affinity(cameras, sub_imgid2cam, pose_mat, num_joints):
    """Serial all-pairs form, float32 stores (src/utils/matching.py:93-113)."""
    M = len(pose_mat)
    pose_mat = np.array(pose_mat)
    D = np.zeros((M, M, num_joints), dtype=np.float32)
    A = np.ones((M, M), dtype=np.float32) * 25
    np.fill_diagonal(A, 0)
    for i in range(M - 1):
        for j in range(i + 1, M):
            ci, cj = sub_imgid2cam[i], sub_imgid2cam[j]
            if ci == cj:
                continue
            dist = epipolar_distance(cameras[ci], pose_mat[i], cameras[cj], pose_mat[j])
            sym = [(d[0] + d[1]) / 2 for d in dist]
            D[i, j] = sym
            D[j, i] = sym
            A[i, j] = A[j, i] = np.mean(sym)
    return A, D


def epipolar_affinity_parallel(cameras, sub_imgid2cam, pose_mat, num_joints):
    """Vectorised per-track form, float64 (src/utils/matching.py:115-151)."""
    M = len(pose_mat)
    pose_mat = np.array(pose_mat)
    homo = np.concatenate([np.flip(pose_mat[:, :, :2], axis=2), np.ones((M, num_joints, 1))], axis=2)
    src = np.transpose(np.repeat(homo, M, 0), (0, 2, 1))
    dst = np.tile(homo, (M, 1, 1))
    Fs = []
    for i in range(M):
        for j in range(M):
            ci, cj = sub_imgid2cam[i], sub_imgid2cam[j]
            Fs.append(np.zeros((3, 3)) if ci == cj else cameras[ci].F[cameras[cj].cid].T)
    Fs = np.array(Fs)
    lines = np.transpose(Fs @ src, (0, 2, 1))
    nu = la.norm(lines[:, :, :2], axis=2).reshape(-1, num_joints, 1)
    nu[nu == 0] = 1
    lines /= nu
    nn = np.sum(lines[:, :, :2] ** 2, axis=2)
    nn[nn == 0] = 1
    d = np.abs(np.sum(dst * lines, axis=2)) / np.sqrt(nn)
    D = d.reshape(M, M, -1)
    D = (D + np.transpose(D, (1, 0, 2))) / 2
    return np.mean(D, axis=2), D


def greedy_view_filter(cameras, pose_mat=None, affinity_mat=None, costs=None, next_pose=None, mode="update"):
    """``Greedy_matching`` (src/utils/matching.py:243-295): walk conflicting view pairs
    (affinity < 0, upper triangle, row-major) and drop one view of each still-live pair."""
    n = affinity_mat.shape[0]
    alive = np.arange(n)
    keep = np.ones(n * 2, dtype=int)
    rows, cols = np.where(np.triu(affinity_mat) < 0)
    ray_d = np.zeros(n)
    for r, c in zip(rows, cols):
        if r not in alive or c not in alive:
            continue
        if mode == "update":
            for k in (r, c):
                if ray_d[k] == 0:
                    cam = cameras[k]
                    px = np.array([np.flip(pose_mat[k, 0, :2])])
                    ray = pixel_rays(cam.RK_INV, cam.position, px)
                    ray_d[k] = ray_point_distance(cam.position, ray, np.array([next_pose]))[0]
            drop = r if ray_d[r] > ray_d[c] else c
        else:
            drop = c if np.sum(affinity_mat[r]) > np.sum(affinity_mat[c]) else r
        alive = alive[alive != drop]
        keep[drop * 2:drop * 2 + 2] = 0
    return alive, keep, affinity_mat


# ----------------------------------------------------------------------------------------------
# triangulation (src/utils/construction.py)
# ----------------------------------------------------------------------------------------------


def _dlt_rows(cameras, Ts, pose_mat, lambda_t):
    """(J, 2*Vt, 4) weighted unit rows, view-major (src/utils/construction.py:90-99)."""
    blocks = []
    for cam, pose, T in zip(cameras, pose_mat, Ts):
        uv = np.flip(pose[:, :2], axis=1)
        C = np.multiply(uv.reshape(-1, 1), np.repeat([cam.P[2]], uv.size, 0))
        C = C - np.tile(cam.P[:2], (len(uv), 1))
        C = C / la.norm(C, axis=1).reshape(-1, 1)
        W = np.repeat([exp(-lambda_t * T)], uv.size, 0)
        blocks.append(np.multiply(W.reshape(-1, 1), C).reshape(-1, 2, 4))
    return np.concatenate(blocks, axis=1)


def dlt_joint_filtered(cameras, Ts, pose_mat, lambda_t, remains, joints_views, next_pose=None):
    """``SVD_pose_kernel_jf`` (src/utils/construction.py:89-114), output sized (J, 3)."""
    A = _dlt_rows(cameras, Ts, pose_mat, lambda_t)
    pose3d = np.zeros((A.shape[0], 3))
    keep = remains == 1
    for k, joints in enumerate(joints_views):
        if len(joints) == 0:
            continue
        if k == 0:
            pose3d[joints] = next_pose[joints]
        else:
            Ak = A[joints][keep[joints]].reshape(len(joints), -1, 4)
            _, _, VT = la.svd(Ak)
            X = np.transpose(VT, (0, 2, 1))[:, :, -1]
            pose3d[joints] = X[:, :3] / X[:, 3].reshape(-1, 1)
    return pose3d


def dlt_all_views(cameras, Ts, pose_mat, lambda_t):
    """``SVD_pose_kernel_parallel`` (src/utils/construction.py:116-131)."""
    A = _dlt_rows(cameras, Ts, pose_mat, lambda_t)
    _, _, VT = la.svd(A)
    X = np.transpose(VT, (0, 2, 1))[:, :, -1]
    return X[:, :3] / X[:, 3].reshape(-1, 1)


def dlt_per_joint(cameras, Ts, joints, remains, lambda_t, next_pose=None):
    """``SVD_pose_kernel`` (src/utils/construction.py:64-87), the older per-joint loop."""
    out = []
    for jid, (joint, remain) in enumerate(zip(joints, remains)):
        if len(remain) <= 1:
            out.append(np.array([None, None, None]) if next_pose is None else next_pose[jid])
            continue
        C, W = [], []
        for i, v in enumerate(remain):
            x, y = joint[v][0], joint[v][1]
            Pm = cameras[v].P
            C.append(y * Pm[2] - Pm[0])
            W.append(exp(-lambda_t * Ts[v]) / la.norm(C[2 * i]))
            C.append(x * Pm[2] - Pm[1])
            W.append(exp(-lambda_t * Ts[v]) / la.norm(C[2 * i + 1]))
        A = np.multiply(np.array(W).reshape(-1, 1), C)
        _, _, VT = la.svd(A)
        X = np.transpose(np.transpose(VT)[:, -1])
        X /= X[3]
        out.append(X[:3])
    return out


# ----------------------------------------------------------------------------------------------
# hypotheses (src/tracking/hypothesis.py)
# ----------------------------------------------------------------------------------------------


class Hypothesis:
    def __init__(self, cam, pts, epi_threshold=40):
        self.joints = len(pts)
        self.poses = [pts]
        self.cams = [cam]
        self.threshold = epi_threshold

    def size(self):
        return len(self.poses)

    def merge(self, o_cam, o_pose):
        self.cams.append(o_cam)
        self.poses.append(o_pose)

    def calculate_cost(self, o_cam, o_pose):
        """Confidence-weighted symmetric epipolar cost + veto (src/tracking/hypothesis.py:53-68)."""
        veto = False
        total = 0
        for person, cam in zip(self.poses, self.cams):
            dist = epipolar_distance(cam, person, o_cam, o_pose)
            c = np.mean([(d[0] * a[2] + d[1] * b[2]) / 2 for d, a, b in zip(dist, person, o_pose)]) / self.threshold
            total += c
            if c > 1 and mean_confidence(o_pose) > 0.5:
                veto = True
        return total / len(self.poses), veto

    def get_3dpose_jf(self, init_threshold, lambda_t):
        """First triangulation of a hypothesis (src/tracking/hypothesis.py:23-44)."""
        Ts = [0 for _ in self.poses]
        _, D = epipolar_affinity(self.cams, np.arange(len(self.cams)), self.poses, num_joints=self.joints)
        A = 1 - D / init_threshold
        joints_views = [[] for _ in self.cams]
        keep = np.ones((self.joints, len(self.cams) * 2), dtype=int)
        for j in range(self.joints):
            alive, keep[j], _ = greedy_view_filter(self.cams, affinity_mat=A[:, :, j], mode="init")
            joints_views[len(alive) - 1].append(j)
            if len(alive) < 2:
                return [], [], [], [], False
        pose3d = dlt_joint_filtered(self.cams, Ts, self.poses, lambda_t, keep, joints_views)
        return self.cams, self.poses, pose3d, joints_views, True


# ----------------------------------------------------------------------------------------------
# tracks (src/tracking/IterativeTracker.py:182-395)
# ----------------------------------------------------------------------------------------------


class Track:
    def __init__(self, track_id, time, cameras, poses2d, pose3d, joints_views, args, arm_joints):
        self.track_id = track_id
        self.hits = 1
        self.age = 1
        self.time_since_update = 0
        self.already_update = False
        self.joints = len(pose3d)
        self.poses2d = {cam.cid: {"time": time, "camera": cam, "pose": pose} for cam, pose in zip(cameras, poses2d)}
        self.poses3d = [{"time": time, "pose3d": np.array(pose3d), "joints_views": joints_views}]
        self.velocity_3d = np.array([[0., 0., 0.] for _ in range(self.joints)])
        self.state = TENTATIVE
        self.args = args
        self.arm = list(arm_joints)
        self.not_arm = [j for j in range(self.joints) if j not in self.arm]
        self.last_keep = None     # trace: (J, 2*Vt) keep mask of the last get_3dpose
        self.last_views = None    # trace: camera ids of the views used, in dict order

    def is_tentative(self):
        return self.state == TENTATIVE

    def is_confirmed(self):
        return self.state == CONFIRMED

    def is_deleted(self):
        return self.state == DELETED

    def add_age(self):
        self.already_update = False
        self.age += 1
        self.time_since_update += 1

    def add_pose(self, camera, time, pose):
        self.already_update = True
        self.poses2d.setdefault(camera.cid, dict())
        self.poses2d[camera.cid] = {"time": time, "camera": camera, "pose": pose}

    def update(self, time):
        if self.update_3dpose(time):
            self.update_motion(time)
            self.hits += 1
            self.time_since_update = 0
            if self.state == TENTATIVE and self.hits >= self.args["n_init"]:
                self.state = CONFIRMED
        else:
            self.mark_missed()

    def mark_missed(self):
        if self.state == TENTATIVE and not self.already_update:
            self.state = DELETED
        elif self.time_since_update >= self.args["max_age"]:
            self.state = DELETED

    def update_3dpose(self, time):
        if not self.already_update:
            return False
        Ts, cams, pose_mat = [], [], []
        for v in self.poses2d.values():
            age = time - v["time"]
            if age <= 3:
                Ts.append(age)
                cams.append(v["camera"])
                pose_mat.append(v["pose"])
        if len(cams) < 2:
            return False
        pose3d, joints_views, ok = self.get_3dpose(time, cams, Ts, np.array(pose_mat))
        if not ok:
            return False
        pose3d = self.smooth_3dpose(time, np.array(pose3d))
        self.poses3d.append({"time": time, "pose3d": pose3d, "joints_views": joints_views})
        if time - self.poses3d[0]["time"] > self.args["max_age"]:
            del self.poses3d[0]
        return True

    def get_3dpose(self, time, cameras, Ts, pose_mat):
        last = self.poses3d[-1]
        nxt = last["pose3d"] + self.velocity_3d * (time - last["time"])
        _, D = epipolar_affinity_parallel(cameras, np.arange(len(cameras)), pose_mat, num_joints=self.joints)
        A = 1 - D / self.args["joint_threshold"]
        fail = 0
        joints_views = [[] for _ in cameras]
        keep = np.ones((self.joints, len(cameras) * 2), dtype=int)
        for j, pose in enumerate(np.transpose(pose_mat, (1, 0, 2))):
            alive, keep[j], _ = greedy_view_filter(cameras, pose_mat=pose.reshape(-1, 1, 3),
                                                   affinity_mat=A[:, :, j], next_pose=nxt[j])
            joints_views[len(alive) - 1].append(j)
            if len(alive) < 2:
                fail += 1
        self.last_keep, self.last_views = keep.copy(), [c.cid for c in cameras]
        pose3d = dlt_joint_filtered(cameras, Ts, pose_mat, self.args["lambda_t"], keep, joints_views, nxt)
        return pose3d, joints_views, False if fail > self.joints / 3 else True

    def smooth_3dpose(self, time, pose3d):
        """Gaussian over the time axis, last sample kept (src/tracking/IterativeTracker.py:371-383)."""
        series = np.array([p["pose3d"] for p in self.poses3d] + [pose3d])
        pose3d[self.not_arm] = gaussian_filter1d(series[:, self.not_arm, :].T, sigma=self.args["sigma"],
                                                 mode="reflect")[:, :, -1].T
        if len(self.arm):
            pose3d[self.arm] = gaussian_filter1d(series[:, self.arm, :].T, sigma=self.args["arm_sigma"],
                                                 mode="reflect")[:, :, -1].T
        return pose3d

    def update_motion(self, time):
        """float32 mean of the last <= 5 pose differences (src/tracking/IterativeTracker.py:385-395)."""
        if len(self.poses3d) < 2:
            return
        diffs = []
        for i in range(len(self.poses3d) - 1, 0, -1):
            diffs.append(self.poses3d[i]["pose3d"].astype(np.float32) - self.poses3d[i - 1]["pose3d"].astype(np.float32))
            if len(diffs) > 4:
                break
        self.velocity_3d = np.mean(diffs, axis=0)


# ----------------------------------------------------------------------------------------------
# tracker (src/tracking/IterativeTracker.py:34-180)
# ----------------------------------------------------------------------------------------------


class Tracker:
    """J-parametrised ``IterativeTracker``.

    ``params``: the 16 fields of src/ivclabpose.py:140-156 (dict or attribute object).
    ``arm_joints``: joints smoothed with ``arm_sigma`` (reference: wrists ``[9, 10]`` of COCO-17).
    ``min_valid_joints``: the hard-coded ``> 10`` of src/tracking/IterativeTracker.py:145."""

    def __init__(self, params, arm_joints: Sequence[int] = (9, 10), min_valid_joints: int = 10):
        self.args = dict(params)
        self.arm_joints = tuple(arm_joints)
        self.min_valid_joints = min_valid_joints
        self.unmatched = dict()
        self.tracks: List[Track] = []
        self.tracks_ids = set()
        self.last_assoc = None   # trace: per camera (m,) matched track_id or -1

    def track_restart(self):
        self.unmatched = dict()
        self.tracks = []
        self.tracks_ids = set()

    # -- new-track initialisation (src/tracking/IterativeTracker.py:52-113) --------------------
    def init_target_GD(self, time):
        if len(self.unmatched) < 2:
            return
        for value in self.unmatched.values():
            value["detections"] = np.array([d for d in value["detections"]
                                            if mean_confidence(d) > self.args["conf_threshold"]])
        H: List[Hypothesis] = []
        thr = self.args["epi_threshold"]
        for idx, value in enumerate(self.unmatched.values()):
            cam, dets = value["camera"], value["detections"]
            if idx == 0:
                H = [Hypothesis(cam, d, thr) for d in dets]
                continue
            C = np.zeros((len(H), len(dets)))
            veto_mask = np.zeros_like(C).astype("int32")
            for h, hyp in enumerate(H):
                for p, d in enumerate(dets):
                    C[h, p], veto = hyp.calculate_cost(cam, d)
                    if veto:
                        veto_mask[h, p] = 1
            rows, cols = linear_sum_assignment(C)
            handled = set()
            for h, p in zip(rows, cols):
                handled.add(p)
                if veto_mask[h, p] == 1:
                    H.append(Hypothesis(cam, dets[p], thr))
                else:
                    H[h].merge(cam, dets[p])
            for p, d in enumerate(dets):
                if p not in handled:
                    H.append(Hypothesis(cam, d, thr))
        for hyp in H:
            if hyp.size() > 1:
                cams, poses2d, pose3d, joints_views, ok = hyp.get_3dpose_jf(self.args["init_threshold"],
                                                                             self.args["lambda_t"])
                if not ok:
                    continue
                tid = 0 if len(self.tracks_ids) == 0 else max(self.tracks_ids) + 1
                self.tracks.append(Track(tid, time, cams, poses2d, pose3d, joints_views, self.args, self.arm_joints))
                self.tracks_ids.add(tid)

    # -- one frame (src/tracking/IterativeTracker.py:115-180) -----------------------------------
    def tracking(self, frame_id, camera_list, frame_list, boxes_list, detections_list, build3D="SVD"):
        assert build3D == "SVD", "Please modify BUILD3D to SVD when PERSON_MATCHER == Iterative"
        last_poses, gaps = [], []
        for tr in self.tracks:
            tr.add_age()
            last_poses.append(tr.poses3d[-1]["pose3d"])
            gaps.append(frame_id - tr.poses3d[-1]["time"])
        a2d, lam = self.args["alpha2d"], self.args["lambda_a"]
        self.last_assoc = []
        for camera, boxes, detections in zip(camera_list, boxes_list, detections_list):
            n, m = len(self.tracks), len(detections)
            assoc = np.full(m, -1, dtype=np.int64)
            if n > 0 and m > 0:
                reproj = camera.project_tracks(np.array(last_poses))
                detections = np.array(detections)
                a = np.repeat(reproj, m, 0)
                b = np.tile(detections[:, :, :2], (n, 1, 1))
                c2d = la.norm(a - b, axis=2).reshape(n, m, -1)
                c2d = 1 - np.transpose(c2d.T / (a2d * np.array(gaps)))
                enough = np.sum(c2d > 0, axis=2) > self.min_valid_joints
                with np.errstate(invalid="ignore", divide="ignore"):
                    aff = np.sum(c2d, where=c2d > 0, axis=2) / np.sum(c2d > 0, axis=2)
                aff[~enough] = 0
                aff = np.transpose(aff.T / np.exp(lam * np.array(gaps)))
                aff[np.isnan(aff)] = 0
                rows, cols = linear_sum_assignment(-aff)
                handled = set()
                for ti, di in zip(rows, cols):
                    if aff[ti, di] > 0:
                        self.tracks[ti].add_pose(camera, frame_id, detections[di])
                        handled.add(di)
                        assoc[di] = self.tracks[ti].track_id
                detections = np.delete(detections, list(handled), axis=0)
                boxes = np.delete(boxes, list(handled), axis=0)
            self.last_assoc.append(assoc)
            self.unmatched[camera.cid] = {"camera": camera, "time": frame_id, "bboxes": boxes, "detections": detections}
        for tr in self.tracks:
            tr.update(frame_id)
        self.init_target_GD(frame_id)
        self.tracks = [tr for tr in self.tracks if not tr.is_deleted()]
        return 0.0, 0.0, 0.0

    # -- output contract (src/ivclabpose.py:259-287) -------------------------------------------
    def frame_output(self):
        """``(ids (n_out,), joints (n_out, J, 3), views (n_out, J))`` for tracks that are Confirmed
        and were updated this frame; ``views[k, j]`` = number of views joint j was built from."""
        ids, joints, views = [], [], []
        for tr in self.tracks:
            if tr.time_since_update > 0 or not tr.is_confirmed():
                continue
            ids.append(tr.track_id)
            last = tr.poses3d[-1]
            joints.append(last["pose3d"])
            nv = np.zeros(tr.joints, dtype=np.int32)
            for k, js in enumerate(last["joints_views"]):
                for j in js:
                    nv[j] = k + 1
            views.append(nv)
        J = self.args["num_joints"]
        return (np.array(ids, dtype=np.int32), np.array(joints).reshape(-1, J, 3),
                np.array(views, dtype=np.int32).reshape(-1, J))


def run_stream(stream, params, arm_joints, min_valid_joints=10, cameras=None, T=None, trace=False):
    """Run the oracle over a ``synth.Stream``; returns per-frame ``(ids, joints, views)`` lists
    (and, with ``trace``, the per-frame association decisions)."""
    if cameras is None:
        cameras = build_cameras(stream.rig["P"], stream.rig["K"], stream.rig["RT"],
                                stream.rig.get("width", 640), stream.rig.get("height", 480))
    trk = Tracker(params, arm_joints, min_valid_joints)
    out, assoc = [], []
    for t in range(stream.T if T is None else T):
        trk.tracking(t, cameras, [None] * len(cameras), stream.frame_boxes(t), stream.frame_detections(t), "SVD")
        out.append(trk.frame_output())
        if trace:
            assoc.append([a.copy() for a in trk.last_assoc])
    return (out, assoc, trk) if trace else out
