"""Stateless batched geometry ops on the GPU (numpy in / numpy out, or torch CUDA tensors).

Each function mirrors one reference function of the hot path and calls the
corresponding ``pam_*`` entry point of ``libpam.so`` (``include/pam.h``).  The camera constants of
the camera list in use are uploaded once and reused while the same list is passed again."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _capi, camera as _camera
from .tracker import PamError, _check

_DEFAULT_PARAMS = dict(conf_threshold=0.5, epi_threshold=60, init_threshold=30, joint_threshold=60, n_init=3,
                       max_age=10, alpha2d=70, lambda_a=3, lambda_t=5, sigma=0.3, arm_sigma=0.8)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("pam_b200.ops needs a CUDA device: there is no CPU fallback")
    return torch


class GeometryOps:
    """One ``pam_handle`` bound to a camera list and a joint count."""

    def __init__(self, cameras: Sequence, num_joints: int, params: Optional[dict] = None, device: int = 0,
                 min_valid_joints: int = 10):
        torch = _torch()
        self.lib = _capi.load_library()
        p = dict(_DEFAULT_PARAMS)
        if params:
            g = (lambda k: params[k]) if isinstance(params, dict) else (lambda k: getattr(params, k))
            for k in list(p):
                try:
                    p[k] = g(k)
                except (KeyError, AttributeError):
                    pass
        p["num_joints"] = int(num_joints)
        self.V, self.J, self.device = len(cameras), int(num_joints), int(device)
        self.cfg = _capi.make_config(p, self.V, 1, 1, (), min_valid_joints)
        self.handle = C.c_void_p()
        _check(self.lib, None, self.lib.pam_create(C.byref(self.cfg), self.device, C.byref(self.handle)))
        self.dev = torch.device(f"cuda:{self.device}")
        self._cam_key = None
        self.set_cameras(cameras)

    def set_cameras(self, cameras):
        key = tuple(id(c) for c in cameras)
        if key == self._cam_key:
            return
        assert len(cameras) == self.V
        arrs = _camera.pack_cameras(cameras)
        _check(self.lib, self.handle,
               self.lib.pam_set_cameras(self.handle, *[C.c_void_p(a.ctypes.data) for a in arrs]))
        self._cam_key, self._cams = key, list(cameras)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.pam_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -------------------------------------------------------------------------------
    def _dev(self, a, dtype):
        torch = _torch()
        if isinstance(a, torch.Tensor):
            return a.to(device=self.dev, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.dev)

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _call(self, fn, *args):
        _check(self.lib, self.handle, fn(self.handle, *args))

    # -- a2 --------------------------------------------------------------------------------------
    def project_points(self, points3d, as_numpy=True):
        """(n, J, 3) -> (V, n, J, 2) as (v, u)."""
        torch = _torch()
        X = self._dev(points3d, torch.float64)
        n_pts = X.numel() // 3
        out = torch.empty((self.V,) + tuple(X.shape[:-1]) + (2,), dtype=torch.float64, device=self.dev)
        self._call(self.lib.pam_project_points, self._p(X), n_pts, self._p(out), self._stream())
        return out.cpu().numpy() if as_numpy else out

    # -- a3 --------------------------------------------------------------------------------------
    def assoc_affinity(self, tracks3d, dt, detections_list, as_numpy=True):
        """tracks (n,J,3), dt (n,), per-camera detection arrays (m_c,J,3) -> list of (n, m_c)."""
        torch = _torch()
        n = len(tracks3d)
        counts = np.array([len(d) for d in detections_list], np.int32)
        mmax = int(max(1, counts.max(initial=0)))
        dets = np.zeros((self.V, mmax, self.J, 3))
        for c, d in enumerate(detections_list):
            if len(d):
                dets[c, :len(d)] = np.asarray(d, dtype=np.float64)
        X = self._dev(np.asarray(tracks3d, dtype=np.float64), torch.float64)
        aff = torch.empty((self.V, n, mmax), dtype=torch.float64, device=self.dev)
        d_dt, d_dets, d_counts = self._dev(np.asarray(dt), torch.int32), self._dev(dets, torch.float64), self._dev(counts, torch.int32)
        self._call(self.lib.pam_assoc_affinity, self._p(X), self._p(d_dt), self._p(d_dets), self._p(d_counts), n, mmax,
                   self._p(aff), self._stream())
        if not as_numpy:
            return aff, counts
        a = aff.cpu().numpy()
        return [a[c, :, :counts[c]] for c in range(self.V)]

    # -- a4 --------------------------------------------------------------------------------------
    def assign(self, cost, maximize=False):
        """``scipy.optimize.linear_sum_assignment`` semantics: (rows, cols) for one (nr, nc) matrix,
        or a list of such pairs for a (B, nr, nc) batch."""
        torch = _torch()
        cost = np.asarray(cost, dtype=np.float64)
        single = cost.ndim == 2
        cb = cost[None] if single else cost
        B, nr, nc = cb.shape
        out = torch.full((B, max(nr, 1)), -1, dtype=torch.int32, device=self.dev)
        if nr and nc:
            d_cost = self._dev(cb, torch.float64)
            self._call(self.lib.pam_assign, self._p(d_cost), B, nr, nc, 1 if maximize else 0, self._p(out), self._stream())
        o = out.cpu().numpy()
        res = []
        for b in range(B):
            rows = np.nonzero(o[b, :nr] >= 0)[0]
            res.append((rows.astype(np.int64), o[b, rows].astype(np.int64)))
        return res[0] if single else res

    # -- a6 / a7 ---------------------------------------------------------------------------------
    def epipolar_pairs(self, cam_index, pose_mat, as_numpy=True):
        """``epipolar_affinity_parallel``: -> (mean (M,M) f64, D (M,M,J) f64)."""
        torch = _torch()
        pose = self._dev(np.asarray(pose_mat, dtype=np.float64), torch.float64)
        M = pose.shape[0]
        D = torch.empty((M, M, self.J), dtype=torch.float64, device=self.dev)
        mean = torch.empty((M, M), dtype=torch.float64, device=self.dev)
        d_cam = self._dev(np.asarray(cam_index), torch.int32)
        self._call(self.lib.pam_epipolar_pairs, self._p(pose), self._p(d_cam), M, self._p(D), self._p(mean), self._stream())
        return (mean.cpu().numpy(), D.cpu().numpy()) if as_numpy else (mean, D)

    def epipolar_allpairs(self, cam_index, pose_mat, want_dist=True, as_numpy=True):
        """``epipolar_affinity``: -> (aff (M,M) f32, D (M,M,J) f32 or None)."""
        torch = _torch()
        pose = pose_mat if isinstance(pose_mat, torch.Tensor) else np.asarray(pose_mat, dtype=np.float64)
        pose = self._dev(pose, torch.float64)
        M = pose.shape[0]
        aff = torch.empty((M, M), dtype=torch.float32, device=self.dev)
        D = torch.empty((M, M, self.J), dtype=torch.float32, device=self.dev) if want_dist else None
        d_cam = self._dev(cam_index if isinstance(cam_index, torch.Tensor) else np.asarray(cam_index), torch.int32)
        self._call(self.lib.pam_epipolar_allpairs, self._p(pose), self._p(d_cam), M, self._p(aff), self._p(D), self._stream())
        if not as_numpy:
            return aff, D
        return aff.cpu().numpy(), (D.cpu().numpy() if D is not None else None)

    def epipolar_distance(self, cam1, pose1, cam2, pose2):
        """Batched ``epipolar_distance``: pose1/pose2 (B,J,3), cam1/cam2 (B,) -> (B,J,2)."""
        torch = _torch()
        p1 = self._dev(np.asarray(pose1, dtype=np.float64), torch.float64)
        p2 = self._dev(np.asarray(pose2, dtype=np.float64), torch.float64)
        B = p1.shape[0]
        out = torch.empty((B, self.J, 2), dtype=torch.float64, device=self.dev)
        d_c1, d_c2 = self._dev(np.asarray(cam1), torch.int32), self._dev(np.asarray(cam2), torch.int32)
        self._call(self.lib.pam_epipolar_distance, self._p(p1), self._p(p2), self._p(d_c1), self._p(d_c2), B, self._p(out),
                   self._stream())
        return out.cpu().numpy()

    # -- a8 / a17 --------------------------------------------------------------------------------
    def view_filter(self, cam_index, affinity, mode="update", uv=None, next_pose=None):
        """Batched ``Greedy_matching``: affinity (B,n,n); update mode needs uv (B,n,2) as (u,v) and
        next_pose (B,3).  -> keep (B,n) uint8."""
        torch = _torch()
        A = np.asarray(affinity)
        B, n = A.shape[0], A.shape[1]
        keep = torch.empty((B, n), dtype=torch.uint8, device=self.dev)
        cam = self._dev(np.asarray(cam_index), torch.int32)
        if mode == "update":
            d_A = self._dev(A.astype(np.float64), torch.float64)
            d_uv = self._dev(np.asarray(uv, dtype=np.float64), torch.float64)
            d_next = self._dev(np.asarray(next_pose, dtype=np.float64), torch.float64)
            self._call(self.lib.pam_view_filter, 0, self._p(d_A), None, self._p(d_uv), self._p(cam), self._p(d_next), B, n,
                       self._p(keep), self._stream())
        else:
            d_A = self._dev(A.astype(np.float32), torch.float32)
            self._call(self.lib.pam_view_filter, 1, None, self._p(d_A), None, self._p(cam), None, B, n, self._p(keep),
                       self._stream())
        return keep.cpu().numpy()

    # -- a10 / a11 -------------------------------------------------------------------------------
    def triangulate(self, cam_index, pose_mat, weights, keep=None, next_pose=None, as_numpy=True):
        """pose_mat (B,Vt,J,3) or (Vt,J,3); cam_index (B,Vt)/(Vt,); weights (B,Vt)/(Vt,) =
        exp(-lambda_t T); keep (B,J,Vt) or None; next_pose (B,J,3) or None -> (B,J,3) / (J,3)."""
        torch = _torch()
        pose = pose_mat if isinstance(pose_mat, torch.Tensor) else np.asarray(pose_mat, dtype=np.float64)
        single = pose.ndim == 3
        pose = self._dev(pose, torch.float64)
        if single:
            pose = pose[None]
        B, Vt = pose.shape[0], pose.shape[1]
        cam = self._dev(np.asarray(cam_index), torch.int32).reshape(B, Vt)
        w = self._dev(np.asarray(weights, dtype=np.float64), torch.float64).reshape(B, Vt)
        kp = None if keep is None else self._dev(np.asarray(keep), torch.uint8).reshape(B, self.J, Vt)
        nx = None if next_pose is None else self._dev(np.asarray(next_pose, dtype=np.float64), torch.float64).reshape(B, self.J, 3)
        out = torch.empty((B, self.J, 3), dtype=torch.float64, device=self.dev)
        self._call(self.lib.pam_triangulate, self._p(pose.contiguous()), self._p(cam), self._p(w), self._p(kp), self._p(nx),
                   B, Vt, self._p(out), self._stream())
        if single:
            out = out[0]
        return out.cpu().numpy() if as_numpy else out

    # -- f4: alternatives without a live caller in the reference ----------------------------------
    def top_down(self, cam_index, poses2d, want_errors=False):
        """``top_down_pose_kernel`` (src/utils/construction.py:9-31): poses2d (B,Vt,J,2) or (Vt,J,2) as (x, y);
        cam_index (B,Vt)/(Vt,) -> pose3d (B,J,3)/(J,3), winning pair (B,2)/(2,) [, errors (B, Vt(Vt-1)/2)]."""
        torch = _torch()
        p = np.asarray(poses2d, dtype=np.float64)
        single = p.ndim == 3
        p = self._dev(p[None] if single else p, torch.float64).contiguous()
        B, Vt = p.shape[0], p.shape[1]
        cam = self._dev(np.asarray(cam_index), torch.int32).reshape(B, Vt)
        out = torch.empty((B, self.J, 3), dtype=torch.float64, device=self.dev)
        pair = torch.empty((B, 2), dtype=torch.int32, device=self.dev)
        err = torch.empty((B, Vt * (Vt - 1) // 2), dtype=torch.float64, device=self.dev) if want_errors else None
        self._call(self.lib.pam_top_down, self._p(p), self._p(cam), B, Vt, self._p(out), self._p(pair), self._p(err),
                   self._stream())
        res = (out.cpu().numpy(), pair.cpu().numpy()) + ((err.cpu().numpy(),) if want_errors else ())
        return tuple(r[0] for r in res) if single else res

    # -- a16 -------------------------------------------------------------------------------------
    def mean_confidence(self, poses):
        """Batched ``get_believe``: (B,J,3) -> (B,)."""
        torch = _torch()
        p = self._dev(np.asarray(poses, dtype=np.float64).reshape(-1, self.J, 3), torch.float64)
        out = torch.empty(p.shape[0], dtype=torch.float64, device=self.dev)
        self._call(self.lib.pam_mean_confidence, self._p(p), p.shape[0], self._p(out), self._stream())
        return out.cpu().numpy()

    def hypothesis_cost(self, hyp_cam, hyp_poses, other_cam: int, other_poses):
        """``Hypothesis.calculate_cost`` of one hypothesis against B detections: -> (cost (B,), veto (B,))."""
        torch = _torch()
        hp = self._dev(np.asarray(hyp_poses, dtype=np.float64), torch.float64)
        hc = self._dev(np.asarray(hyp_cam), torch.int32)
        op = self._dev(np.asarray(other_poses, dtype=np.float64).reshape(-1, self.J, 3), torch.float64)
        B = op.shape[0]
        cost = torch.empty(B, dtype=torch.float64, device=self.dev)
        veto = torch.empty(B, dtype=torch.uint8, device=self.dev)
        self._call(self.lib.pam_hypothesis_cost, self._p(hp), self._p(hc), hp.shape[0], self._p(op), int(other_cam), B,
                   self._p(cost), self._p(veto), self._stream())
        return cost.cpu().numpy(), veto.cpu().numpy().astype(bool)

    # -- a9 --------------------------------------------------------------------------------------
    def ray_distance(self, camera: int, uv, points3d=None, want_dirs=False):
        """uv (n,2) as (u,v) -> (dist (n,) or None, dirs (n,3) or None)."""
        torch = _torch()
        p = self._dev(np.asarray(uv, dtype=np.float64).reshape(-1, 2), torch.float64)
        n = p.shape[0]
        X = None if points3d is None else self._dev(np.asarray(points3d, dtype=np.float64).reshape(-1, 3), torch.float64)
        dist = torch.empty(n, dtype=torch.float64, device=self.dev) if X is not None else None
        dirs = torch.empty((n, 3), dtype=torch.float64, device=self.dev) if want_dirs else None
        self._call(self.lib.pam_ray_distance, int(camera), self._p(p), self._p(X), n, self._p(dist), self._p(dirs),
                   self._stream())
        return (None if dist is None else dist.cpu().numpy()), (None if dirs is None else dirs.cpu().numpy())


# ------------------------------------------------------------------------------------------------
# handle cache for the drop-in modules: one handle per (number of cameras, joint count); the camera
# constants are re-uploaded only when a different camera list shows up.
# ------------------------------------------------------------------------------------------------
_cache = {}


def get_ops(cameras: Sequence, num_joints: int, params: Optional[dict] = None) -> GeometryOps:
    sig = tuple(sorted(params.items())) if isinstance(params, dict) else None
    key = (len(cameras), int(num_joints), sig)
    ops = _cache.get(key)
    if ops is None:
        ops = GeometryOps(cameras, num_joints, params)
        _cache[key] = ops
    else:
        ops.set_cameras(cameras)
    return ops


def project_points(cameras, points3d):
    pts = np.asarray(points3d)
    return get_ops(cameras, pts.shape[-2]).project_points(pts)


# ------------------------------------------------------------------------------------------------
# One-Euro filter bank: needs no cameras, so it runs on a bare handle per device
# ------------------------------------------------------------------------------------------------
_bare = {}


def _bare_handle(device: int = 0):
    if device not in _bare:
        lib = _capi.load_library()
        cfg = _capi.make_config(dict(_DEFAULT_PARAMS, num_joints=1), 1, 1, 1, (), 0)
        h = C.c_void_p()
        _check(lib, None, lib.pam_create(C.byref(cfg), int(device), C.byref(h)))
        _bare[device] = (lib, h)
    return _bare[device]


def one_euro(x, state, freq, mincutoff, beta, dcutoff, device: int = 0):
    """One step of a One-Euro filter bank (src/tracking/OneEuroFilter.py:60-77): ``x`` (n,) float64, ``state`` a CUDA
    tensor (n, 4) float64 (zeros = fresh, updated in place) -> filtered values (n,) as numpy."""
    torch = _torch()
    lib, h = _bare_handle(device)
    xd = torch.as_tensor(np.asarray(x, dtype=np.float64).reshape(-1), device=f"cuda:{device}")
    out = torch.empty_like(xd)
    st = torch.cuda.current_stream(device).cuda_stream
    rc = lib.pam_one_euro(h, C.c_void_p(xd.data_ptr()), xd.shape[0], float(freq), float(mincutoff), float(beta),
                          float(dcutoff), C.c_void_p(state.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(st))
    _check(lib, h, rc)
    return out.cpu().numpy()


def kalman9(meas, state, hz=25.0, has_meas=None, device: int = 0):
    """One step of a bank of the reference's Kalman smoothers (src/tracking/KalmanFilter.py:52-65, ``predict(pt3d)``):
    ``meas`` (n, 3) float64 or None (predict only), ``state`` a CUDA tensor (n, 90) float32 updated in place
    -> predicted positions (n, 3) as numpy."""
    torch = _torch()
    lib, h = _bare_handle(device)
    n = state.shape[0]
    md = None if meas is None else torch.as_tensor(np.asarray(meas, dtype=np.float64).reshape(n, 3), device=f"cuda:{device}")
    hm = None if has_meas is None else torch.as_tensor(np.asarray(has_meas, dtype=np.uint8).reshape(n), device=f"cuda:{device}")
    out = torch.empty((n, 3), dtype=torch.float64, device=f"cuda:{device}")
    st = torch.cuda.current_stream(device).cuda_stream
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    _check(lib, h, lib.pam_kalman9(h, p(md), p(hm), n, float(hz), p(state), p(out), C.c_void_p(st)))
    return out.cpu().numpy()

