"""Host-side mirror of the reference's camera ingest (one-time set-up).

``Camera`` and ``GetCameraParameters`` keep the reference's names, attributes and dtypes
(src/ivclabpose.py:35-46, 162-181): ``P, K, RT, RK_INV, F`` float32, ``position`` float64,
``F[b]`` = fundamental matrix towards camera ``b`` with ``x_a^T F_ab x_b = 0``, ``x = (u, v, 1)``.
``pack_cameras`` flattens a camera list into the four arrays ``pam_set_cameras`` takes."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


class Camera(object):
    def __init__(self, cid, P, K, RT, F, w=640, h=480, RK_INV=None, position=None):
        self.cid = cid
        self.P = P
        self.K = K
        self.RT = RT
        self.F = F
        self.w = w
        self.h = h
        # R^-1 K^-1 in the dtype of the inputs (float32), centre from the 4x4 inverse (float64);
        # the two keyword arguments carry the values of a device ingest (ingest_on_device)
        self.RK_INV = np.linalg.inv(RT[:, :3]) @ np.linalg.inv(K) if RK_INV is None else RK_INV
        self.position = np.linalg.inv(np.vstack([RT, [0, 0, 0, 1]]))[:3, 3] if position is None else position

    def undistort(self, im):
        return im

    def undistort_points(self, points2d):
        return points2d

    def projectPoints(self, points3d):
        """(n, 3) -> (n, 2) as (v, u) (src/ivclabpose.py:62-89)."""
        pts = np.asarray(points3d, dtype=np.float64)
        out = np.zeros((len(pts), 2))
        for i, X in enumerate(pts):
            a, b, c = self.P @ np.array([X[0], X[1], X[2], 1])
            c = 10e-6 if c == 0 else c
            out[i, 1] = a / c
            out[i, 0] = b / c
        return out

    projectPoints_undist = projectPoints

    def projectPoints_parallel(self, points3d):
        """(n, J, 3) -> (n, J, 2) as (v, u) on the GPU (src/ivclabpose.py:91-98)."""
        from . import ops
        return ops.project_points([self], np.asarray(points3d))[0]


def fundamental_tensor(K: np.ndarray, RT: np.ndarray) -> np.ndarray:
    """(V, V, 3, 3) float32, same float32 torch CPU expression as src/ivclabpose.py:166-177 so the
    constants are bit-identical to the reference's."""
    import torch

    V = len(K)
    Kt = [torch.tensor(k) for k in K]
    Rt = [torch.tensor(rt[:, :3]) for rt in RT]
    Tt = [torch.tensor(rt[:, 3]) for rt in RT]
    F = torch.zeros(V, V, 3, 3)
    for a in range(V):
        Kinv_t = torch.inverse(Kt[a]).t()
        for b in range(V):
            Rab = Rt[a] @ Rt[b].t()
            e = Kt[b] @ Rt[b] @ Rt[a].t() @ (Tt[a] - Rab @ Tt[b])
            ex = torch.tensor([[0, -e[2], e[1]], [e[2], 0, -e[0]], [-e[1], e[0], 0]])
            F[a, b] += Kinv_t @ Rab @ Kt[b].t() @ ex
            if F[a, b].sum() == 0:
                F[a, b] += 1e-12
    return F.numpy()


def ingest_on_device(K: np.ndarray, RT: np.ndarray, device: int = 0):
    """``pam_camera_ingest``: (V,3,3), (V,3,4) float32 -> ``RK_INV (V,3,3) f32, position (V,3) f64,
    F (V,V,3,3) f32`` computed by one kernel launch (961 matrices at V=31).
    Equal to the host ingest up to float32 rounding, see include/pam.h."""
    import torch
    from . import _capi

    lib = _capi.load_library()
    V = len(K)
    dev = torch.device("cuda", device)
    dK = torch.from_numpy(np.ascontiguousarray(K, np.float32)).to(dev)
    dRT = torch.from_numpy(np.ascontiguousarray(RT, np.float32)).to(dev)
    dRK = torch.empty((V, 3, 3), dtype=torch.float32, device=dev)
    dpos = torch.empty((V, 3), dtype=torch.float64, device=dev)
    dF = torch.empty((V, V, 3, 3), dtype=torch.float32, device=dev)
    rc = lib.pam_camera_ingest(device, V, dK.data_ptr(), dRT.data_ptr(), dRK.data_ptr(), dpos.data_ptr(), dF.data_ptr(),
                               torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
        raise RuntimeError("pam_camera_ingest: " + lib.pam_last_error(None).decode())
    return dRK.cpu().numpy(), dpos.cpu().numpy(), dF.cpu().numpy()


def GetCameraParameters(camera_parameter, im_width=640, im_height=480, device=None) -> List[Camera]:
    """``{'P': (V,3,4), 'K': (V,3,3), 'RT': (V,3,4)}`` (a ``camera_parameter.pickle``) -> cameras.
    ``device=None`` (default) evaluates the reference's float32 host expression, bit-identical constants;
    ``device=<cuda index>`` derives them with ``pam_camera_ingest`` (float32-rounding level differences)."""
    P = np.asarray(camera_parameter["P"]).astype(np.float32)
    K = np.asarray(camera_parameter["K"]).astype(np.float32)
    RT = np.asarray(camera_parameter["RT"]).astype(np.float32)
    if device is not None:
        RK, pos, F = ingest_on_device(K, RT, device)
        return [Camera(j, P[j], K[j], RT[j], F[j], w=im_width, h=im_height, RK_INV=RK[j], position=pos[j])
                for j in range(len(P))]
    F = fundamental_tensor(K, RT)
    return [Camera(j, P[j], K[j], RT[j], F[j], w=im_width, h=im_height) for j in range(len(P))]


def pack_cameras(cameras: Sequence) -> tuple:
    """-> ``P (V,12) f32, RKinv (V,9) f32, position (V,3) f64, F (V,V,9) f32`` (C-contiguous).
    Works for any object with the reference's ``Camera`` attributes; ``F`` rows are indexed by the
    other camera's ``cid`` like the reference does (src/utils/matching.py:59,136)."""
    V = len(cameras)
    P = np.ascontiguousarray(np.stack([np.asarray(c.P, np.float32).reshape(12) for c in cameras]))
    RK = np.ascontiguousarray(np.stack([np.asarray(c.RK_INV, np.float32).reshape(9) for c in cameras]))
    pos = np.ascontiguousarray(np.stack([np.asarray(c.position, np.float64).reshape(3) for c in cameras]))
    F = np.zeros((V, V, 9), np.float32)
    for a, ca in enumerate(cameras):
        for b, cb in enumerate(cameras):
            F[a, b] = np.asarray(ca.F[cb.cid], np.float32).reshape(9)
    return P, RK, pos, np.ascontiguousarray(F)
