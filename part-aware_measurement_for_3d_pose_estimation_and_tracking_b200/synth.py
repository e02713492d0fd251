"""Seeded synthetic multi-view streams.

The CNN front-end of the reference (YOLOv3 + HRNet, src/ivclabpose.py:183-214) is out of
scope; this module stands in for it.  It builds a ring of pin-hole cameras, lets ``P``
articulated skeletons walk on smooth ground-plane paths, projects them through the
camera models and emits noisy, permuted, occasionally missing / outlier-ridden 2-D
detections in the reference's own layout: per camera an array ``(m, J, 3)`` of
``(v, u, conf)`` = (row, col, confidence)  (src/ivclabpose.py:238-244).

Everything is deterministic from ``numpy.random.default_rng(1000 * config_id + seq_id)``.
Detections are rounded to float32 so that the float64 CPU oracle and the float32 device
buffers see *identical* numbers.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------------------------
# shapes named by BASELINE.json "configs"
# ----------------------------------------------------------------------------------------------


@dataclass(frozen=True)
class Shape:
    name: str
    config_id: int
    V: int            # cameras
    P: int            # people
    J: int            # joints
    T: int            # frames
    width: int
    height: int
    focal: float
    radius: float     # camera ring radius (m)
    cam_height: float
    arena: float      # radius of the circle the people's "home" points sit on (m)
    orbit: float      # radius of each person's own loop around the home point (m)
    arm_joints: tuple  # wrists: smoothed with ARM_SIGMA (src/tracking/IterativeTracker.py:382)


SHAPES: Dict[str, Shape] = {
    # 0: "Campus-shaped synthetic: 3 cameras, 3 people, 14 joints, 2000 frames"
    "campus": Shape("campus", 0, 3, 3, 14, 2000, 360, 288, 440.0, 8.0, 3.0, 1.2, 0.45, (6, 11)),
    # 1: "Shelf-shaped synthetic: 5 cameras, 4 people, 14 joints, 3200 frames"
    "shelf": Shape("shelf", 1, 5, 4, 14, 3200, 1032, 776, 1060.0, 4.5, 2.4, 1.1, 0.40, (6, 11)),
    # 2: "Panoptic-shaped synthetic: 5 HD cameras, 8 people, 19 COCO joints, 10k frames"
    "panoptic": Shape("panoptic", 2, 5, 8, 19, 10000, 1920, 1080, 1500.0, 4.2, 2.2, 1.45, 0.20, (5, 11)),
    # 3: "Dense crowd stress: 31 cameras, 64 people, 19 joints per frame"
    "dense": Shape("dense", 3, 31, 64, 19, 8, 1920, 1080, 1300.0, 9.0, 4.0, 4.0, 0.15, (5, 11)),
    # reference-native joint count (the unmodified reference only runs J = 17)
    "shelf17": Shape("shelf17", 5, 5, 4, 17, 400, 1032, 776, 1060.0, 4.5, 2.4, 1.1, 0.40, (9, 10)),
    "campus17": Shape("campus17", 6, 3, 3, 17, 400, 360, 288, 440.0, 8.0, 3.0, 1.2, 0.45, (9, 10)),
}

# tracker hyper-parameters per dataset (src/configs/*/model_configs.yaml, table in SURVEY.md section 5)
TRACKER_PARAMS: Dict[str, dict] = {
    "campus": dict(conf_threshold=0.4, epi_threshold=25, init_threshold=15, joint_threshold=15,
                   n_init=3, max_age=10, alpha2d=30, lambda_a=3, lambda_t=5, sigma=0.6, arm_sigma=0.8),
    "shelf": dict(conf_threshold=0.5, epi_threshold=60, init_threshold=30, joint_threshold=60,
                  n_init=3, max_age=10, alpha2d=70, lambda_a=3, lambda_t=5, sigma=0.3, arm_sigma=0.8),
    "panoptic": dict(conf_threshold=0.4, epi_threshold=60, init_threshold=50, joint_threshold=30,
                     n_init=3, max_age=10, alpha2d=60, lambda_a=3, lambda_t=5, sigma=0.3, arm_sigma=0.8),
}
TRACKER_PARAMS["dense"] = TRACKER_PARAMS["panoptic"]
TRACKER_PARAMS["shelf17"] = TRACKER_PARAMS["shelf"]
TRACKER_PARAMS["campus17"] = TRACKER_PARAMS["campus"]


def tracker_params(shape: "Shape | str") -> dict:
    """The 16 fields ``ivclabpose`` copies into ``iter_args`` (src/ivclabpose.py:140-156)."""
    sh = SHAPES[shape] if isinstance(shape, str) else shape
    p = dict(TRACKER_PARAMS[sh.name])
    p.update(num_joints=sh.J, init_method="GD", w2d=0.4, w3d=0.6, alpha3d=0.15)
    return p


# ----------------------------------------------------------------------------------------------
# skeleton templates: body frame x = right, y = forward, z = up, unit = body height
# ----------------------------------------------------------------------------------------------

_COCO17 = np.array([
    [0.00, 0.06, 0.93],   # 0 nose
    [0.03, 0.05, 0.95],   # 1 l eye
    [-0.03, 0.05, 0.95],  # 2 r eye
    [0.07, 0.00, 0.94],   # 3 l ear
    [-0.07, 0.00, 0.94],  # 4 r ear
    [0.11, 0.00, 0.82],   # 5 l shoulder
    [-0.11, 0.00, 0.82],  # 6 r shoulder
    [0.14, 0.00, 0.64],   # 7 l elbow
    [-0.14, 0.00, 0.64],  # 8 r elbow
    [0.15, 0.03, 0.48],   # 9 l wrist
    [-0.15, 0.03, 0.48],  # 10 r wrist
    [0.06, 0.00, 0.53],   # 11 l hip
    [-0.06, 0.00, 0.53],  # 12 r hip
    [0.07, 0.00, 0.29],   # 13 l knee
    [-0.07, 0.00, 0.29],  # 14 r knee
    [0.07, 0.00, 0.04],   # 15 l ankle
    [-0.07, 0.00, 0.04],  # 16 r ankle
])
# swing group per joint: 0 none, +-1 arm (l/r), +-2 leg (l/r); amplitude grows toward the extremity
_COCO17_SWING = np.array([0, 0, 0, 0, 0, 0, 0, .5, -.5, 1., -1., 0, 0, -.6, .6, -1.2, 1.2])

# Shelf/Campus 14-joint order (src/eval/transformation.py:5-39): r-ankle r-knee r-hip l-hip l-knee
# l-ankle r-wrist r-elbow r-shoulder l-shoulder l-elbow l-wrist bottom-head top-head
_SHELF14_FROM17 = [16, 14, 12, 11, 13, 15, 10, 8, 6, 5, 7, 9]
_SHELF14 = np.vstack([_COCO17[_SHELF14_FROM17], [[0.0, 0.0, 0.86]], [[0.0, 0.0, 1.0]]])
_SHELF14_SWING = np.concatenate([_COCO17_SWING[_SHELF14_FROM17], [0, 0]])

# Panoptic COCO19 order: neck nose bodycentre l-shoulder l-elbow l-wrist l-hip l-knee l-ankle
# r-shoulder r-elbow r-wrist r-hip r-knee r-ankle l-eye l-ear r-eye r-ear
_C19_FROM17 = [None, 0, None, 5, 7, 9, 11, 13, 15, 6, 8, 10, 12, 14, 16, 1, 3, 2, 4]
_COCO19 = np.array([[0.0, 0.0, 0.84] if i is None else _COCO17[i] for i in _C19_FROM17])
_COCO19[2] = [0.0, 0.0, 0.53]
_COCO19_SWING = np.array([0.0 if i is None else _COCO17_SWING[i] for i in _C19_FROM17])


def skeleton_template(J: int):
    if J == 17:
        return _COCO17.copy(), _COCO17_SWING.copy()
    if J == 14:
        return _SHELF14.copy(), _SHELF14_SWING.copy()
    if J == 19:
        return _COCO19.copy(), _COCO19_SWING.copy()
    idx = np.arange(J) % 17
    tpl = _COCO17[idx].copy()
    tpl[:, 2] = np.clip(tpl[:, 2] - 0.02 * (np.arange(J) // 17), 0.02, 1.0)
    return tpl, _COCO17_SWING[idx].copy()


# ----------------------------------------------------------------------------------------------
# rig
# ----------------------------------------------------------------------------------------------


def make_rig(shape: "Shape | str", seed: Optional[int] = None) -> dict:
    """Ring of pin-hole cameras looking at the arena centre.

    Returns the content of a ``camera_parameter.pickle``:  ``{'P': (V,3,4), 'K': (V,3,3),
    'RT': (V,3,4)}`` in float64; ``GetCameraParameters`` casts to float32
    (src/ivclabpose.py:163-165)."""
    sh = SHAPES[shape] if isinstance(shape, str) else shape
    rng = np.random.default_rng(1000 * sh.config_id + 999 if seed is None else seed)
    V = sh.V
    K = np.zeros((V, 3, 3))
    RT = np.zeros((V, 3, 4))
    P = np.zeros((V, 3, 4))
    for i in range(V):
        if V > 8:  # dome: two rings at different heights
            ring = i % 2
            ang = 2 * np.pi * (i + 0.25 * ring) / V + rng.uniform(-0.03, 0.03)
            hgt = sh.cam_height * (1.0 if ring == 0 else 1.6)
        else:
            ang = 2 * np.pi * i / V + rng.uniform(-0.15, 0.15)
            hgt = sh.cam_height + rng.uniform(-0.3, 0.3)
        rad = sh.radius * (1 + rng.uniform(-0.08, 0.08))
        C = np.array([rad * np.cos(ang), rad * np.sin(ang), hgt])
        target = np.array([rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), 0.9])
        z = target - C
        z /= np.linalg.norm(z)
        x = np.cross(z, np.array([0.0, 0.0, 1.0]))
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])
        f = sh.focal * (1 + rng.uniform(-0.03, 0.03))
        K[i] = [[f, 0, sh.width / 2], [0, f, sh.height / 2], [0, 0, 1]]
        RT[i, :, :3] = R
        RT[i, :, 3] = -R @ C
        P[i] = K[i] @ RT[i]
    return dict(P=P, K=K, RT=RT, width=sh.width, height=sh.height)


# ----------------------------------------------------------------------------------------------
# streams
# ----------------------------------------------------------------------------------------------


@dataclass
class Stream:
    shape: Shape
    seq_id: int
    rig: dict
    dets: np.ndarray        # (T, V, D, J, 3) float32, (v, u, conf), zero padded
    counts: np.ndarray      # (T, V) int32
    person_of_det: np.ndarray  # (T, V, D) int32, ground-truth person per detection, -1 = pad
    gt: np.ndarray          # (T, P, J, 3) float64 world joints (metres)
    meta: dict = field(default_factory=dict)

    @property
    def T(self) -> int:
        return self.dets.shape[0]

    def frame_detections(self, t: int) -> List[np.ndarray]:
        """Per camera ``(m, J, 3)`` float64 arrays -- the ``detections_list`` argument of
        ``IterativeTracker.tracking`` (src/tracking/IterativeTracker.py:115)."""
        return [self.dets[t, c, : self.counts[t, c]].astype(np.float64) for c in range(self.dets.shape[1])]

    def frame_boxes(self, t: int) -> List[np.ndarray]:
        return [np.zeros((int(self.counts[t, c]), 4)) for c in range(self.dets.shape[1])]


def make_stream(shape: "Shape | str", seq_id: int = 0, T: Optional[int] = None, *, noise_px: float = 1.0,
                miss_prob: float = 0.02, outlier_prob: float = 0.01, outlier_px: float = 50.0,
                rig: Optional[dict] = None, enter_stagger: int = 0, P: Optional[int] = None,
                absences: Sequence = ()) -> Stream:
    """One seeded sequence of ``T`` frames.

    ``enter_stagger`` > 0 makes person ``p`` appear only from frame ``p * enter_stagger`` on,
    which exercises new-track initialisation (src/tracking/IterativeTracker.py:52-113) in the
    middle of a stream.  ``absences`` = ``[(person, t0, t1), ...]`` removes a person from every
    camera for frames ``t0 <= t < t1`` (track ageing, deletion after ``max_age`` and re-entry under a
    fresh id, src/tracking/IterativeTracker.py:268-274,108-113)."""
    sh = SHAPES[shape] if isinstance(shape, str) else shape
    T = sh.T if T is None else T
    P = sh.P if P is None else P
    V, J = sh.V, sh.J
    rig = make_rig(sh) if rig is None else rig
    rng = np.random.default_rng(1000 * sh.config_id + seq_id)

    tpl, swing = skeleton_template(J)
    height = rng.uniform(1.6, 1.9, size=P)
    home_ang = 2 * np.pi * (np.arange(P) + rng.uniform(-0.1, 0.1, size=P)) / max(P, 1)
    if P > 16:  # dense crowd: homes on a jittered grid inside the arena
        side = int(np.ceil(np.sqrt(P)))
        gx, gy = np.meshgrid(np.arange(side), np.arange(side))
        pitch = 2 * sh.arena / side
        home = (np.stack([gx.ravel(), gy.ravel()], 1)[:P] - (side - 1) / 2) * pitch
    else:
        home = sh.arena * np.stack([np.cos(home_ang), np.sin(home_ang)], 1)
    omega = rng.uniform(0.02, 0.08, size=P) * rng.choice([-1.0, 1.0], size=P)
    phase0 = rng.uniform(0, 2 * np.pi, size=P)
    gait = rng.uniform(0.25, 0.45, size=P)
    gait0 = rng.uniform(0, 2 * np.pi, size=P)

    t = np.arange(T)[:, None]                                   # (T,1)
    th = phase0[None] + omega[None] * t                         # (T,P)
    root = home[None] + sh.orbit * np.stack([np.cos(th), np.sin(th)], -1)   # (T,P,2)
    heading = th + np.sign(omega)[None] * np.pi / 2             # tangent direction
    ch, shd = np.cos(heading), np.sin(heading)
    sw = 0.16 * np.sin(gait0[None] + gait[None] * t)            # (T,P) fore/aft swing (m)
    body = tpl[None, None] * height[None, :, None, None]        # (1,P,J,3)
    bx = np.broadcast_to(body[..., 0], (T, P, J))
    by = body[..., 1] + sw[..., None] * swing[None, None]
    bz = np.broadcast_to(body[..., 2], (T, P, J))
    # body frame -> world: forward axis = (cos h, sin h), right axis = (sin h, -cos h)
    gt = np.empty((T, P, J, 3))
    gt[..., 0] = root[..., 0:1] + by * ch[..., None] + bx * shd[..., None]
    gt[..., 1] = root[..., 1:2] + by * shd[..., None] - bx * ch[..., None]
    gt[..., 2] = bz

    # the same float32 camera constants the tracker will use
    Pm = rig["P"].astype(np.float32).astype(np.float64)         # (V,3,4)
    Xh = np.concatenate([gt, np.ones((T, P, J, 1))], -1)        # (T,P,J,4)
    proj = np.einsum("vik,tpjk->tvpji", Pm, Xh)                 # (T,V,P,J,3)
    uv = proj[..., :2] / proj[..., 2:3]
    uv = uv + rng.normal(0.0, noise_px, size=uv.shape) if noise_px > 0 else uv
    if outlier_prob > 0:
        out = rng.random(size=uv.shape[:-1]) < outlier_prob
        uv = uv + out[..., None] * rng.uniform(-outlier_px, outlier_px, size=uv.shape)
    conf = rng.uniform(0.7, 1.0, size=uv.shape[:-1])
    vuc = np.stack([uv[..., 1], uv[..., 0], conf], -1).astype(np.float32)   # (T,V,P,J,3) (v,u,conf)

    present = rng.random(size=(T, V, P)) >= miss_prob
    if enter_stagger > 0:
        present &= (np.arange(T)[:, None, None] >= (np.arange(P) * enter_stagger)[None, None, :])
    for (p, t0, t1) in absences:
        present[t0:t1, :, p] = False
    order = np.argsort(rng.random(size=(T, V, P)), axis=-1)     # random person order per camera/frame

    D = P
    dets = np.zeros((T, V, D, J, 3), np.float32)
    pod = np.full((T, V, D), -1, np.int32)
    pres_o = np.take_along_axis(present, order, -1)             # presence in emitted order
    slot = np.cumsum(pres_o, -1) - 1                            # output slot of each emitted person
    counts = pres_o.sum(-1).astype(np.int32)
    ti, vi, ki = np.nonzero(pres_o)
    pi = order[ti, vi, ki]
    si = slot[ti, vi, ki]
    dets[ti, vi, si] = vuc[ti, vi, pi]
    pod[ti, vi, si] = pi
    return Stream(sh, seq_id, rig, dets, counts, pod, gt,
                  dict(noise_px=noise_px, miss_prob=miss_prob, outlier_prob=outlier_prob,
                       enter_stagger=enter_stagger, absences=list(absences)))


def make_batch(shape: "Shape | str", n_seq: int, T: Optional[int] = None, seq0: int = 0, **kw):
    """``n_seq`` independent sequences over ONE rig, packed for the device:
    ``dets (S, T, V, D, J, 3) f32``, ``counts (S, T, V) i32``, ``gt (S, T, P, J, 3)``."""
    sh = SHAPES[shape] if isinstance(shape, str) else shape
    rig = make_rig(sh)
    streams = [make_stream(sh, seq0 + s, T, rig=rig, **kw) for s in range(n_seq)]
    return (rig, np.stack([s.dets for s in streams]), np.stack([s.counts for s in streams]),
            np.stack([s.gt for s in streams]), streams)
