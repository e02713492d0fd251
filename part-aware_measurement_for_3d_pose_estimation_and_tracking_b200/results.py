"""Result packing in the reference's formats: the tuple
``PersonTrack_Project3DPose`` returns (src/ivclabpose.py:259-287), the ``{frame: (n, 3, J)}`` pickle of
``Write3DResult`` (src/evalmodel.py:373-377) and the per-camera JSON of ``Write2DResult`` (:352-371).
Host-side glue over the tracker's output tensors."""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, List, Sequence

import numpy as np


def _np(a):
    return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)


def person_track_output(out: dict, s: int, t: int, dets=None, n_views: int = 0):
    """``(camera_ids, pts, person_ids, pts3d, pts3d_joints_views, person3d_ids)`` of frame ``t`` of
    sequence ``s``: ``pts3d (n, 3, J)``, ``person3d_ids (n,)``, ``pts3d_joints_views`` = per track the
    list over k of joints built from k+1 views; with ``dets (S,T,V,D,J,3)`` and an ``assoc`` output also
    the per-track camera ids / 2-D poses matched this frame."""
    k = int(_np(out["count"])[s, t])
    ids = _np(out["ids"])[s, t, :k].astype(np.int64)
    pts3d = np.transpose(_np(out["joints"])[s, t, :k].astype(np.float64), (0, 2, 1))
    views = []
    if out.get("nviews") is not None:
        nv = _np(out["nviews"])[s, t, :k]
        for r in nv:
            jv = [[] for _ in range(max(n_views, int(r.max(initial=0))))]
            for j, c in enumerate(r):
                if c > 0:
                    jv[int(c) - 1].append(j)
            views.append(jv)
    camera_ids, pts, person_ids = [], [], []
    if dets is not None and out.get("assoc") is not None:
        assoc = _np(out["assoc"])[s, t]
        d = _np(dets)[s, t]
        for tid in ids:
            cams = [int(c) for c in range(assoc.shape[0]) if (assoc[c] == tid).any()]
            camera_ids.append(cams)
            pts.append([d[c, int(np.nonzero(assoc[c] == tid)[0][0])].astype(np.float64) for c in cams])
            person_ids.append([int(tid)] * len(cams))
    return (np.array(camera_ids, dtype="object"), np.array(pts, dtype="object"), person_ids, pts3d, views, ids)


def multi_poses3d(out: dict, s: int, frame_ids: Sequence = None) -> Dict:
    """``{frame_id: ndarray (n, 3, J)}`` -- what evalmodel.py:83 collects and Write3DResult pickles."""
    cnt, joints = _np(out["count"])[s], _np(out["joints"])[s]
    T = cnt.shape[0]
    frame_ids = range(T) if frame_ids is None else frame_ids
    return {fid: np.transpose(joints[t, :cnt[t]].astype(np.float64), (0, 2, 1)) for t, fid in enumerate(frame_ids)}


def write_3d_result(poses: Dict, filepath: str):
    """``Write3DResult`` (src/evalmodel.py:373-377)."""
    d = os.path.dirname(filepath)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(filepath, "wb") as f:
        pickle.dump(poses, f)


def write_2d_result(image_wh, annotations: List[dict], save_dir="TrackResult"):
    """``Write2DResult`` (src/evalmodel.py:352-371): one ``Camera<cid>.json`` per camera."""
    os.makedirs(save_dir, exist_ok=True)
    cameras = {}
    for a in annotations:
        cam = "Camera" + str(a["cid"])
        ts = a["timestamp"]
        name = cam + os.sep + str(ts) + ".jpg"
        cameras.setdefault(cam, {"image_wh": [image_wh[1], image_wh[0]], "frames": {}})
        cameras[cam]["frames"].setdefault(name, {"camera": cam, "timestamp": float(ts), "poses": []})
        cameras[cam]["frames"][name]["poses"].append({"id": int(a["pid"]), "points_2d": np.flip(a["pose"], axis=1).tolist(),
                                                      "scores": np.asarray(a["scores"]).tolist()})
    for key, value in cameras.items():
        with open(os.path.join(save_dir, key + ".json"), "w") as fp:
            json.dump(value, fp)
