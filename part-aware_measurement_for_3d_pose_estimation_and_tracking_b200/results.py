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


def joints_views_list(nviews_row, views_used: int):
    """``joints_views`` as ``IterTrack.get_3dpose`` builds it (src/tracking/IterativeTracker.py:350-357): one bucket
    per usable view of the update, bucket ``k - 1`` lists the joints built from ``k`` views -- and, like the
    reference's ``joints_views[len(matched_list) - 1]``, a joint left with 0 views lands in the LAST bucket."""
    jv = [[] for _ in range(int(views_used))]
    if views_used:
        for j, k in enumerate(nviews_row):
            jv[int(k) - 1].append(j)
    return jv


def person_track_output(out: dict, s: int, t: int, dets=None, n_views: int = 0):
    """``(camera_ids, pts, person_ids, pts3d, pts3d_joints_views, person3d_ids)`` of frame ``t`` of sequence ``s``
    exactly as ``PersonTrack_Project3DPose`` returns them (src/ivclabpose.py:259-287): ``pts3d (n, 3, J)``,
    ``person3d_ids (n,)``, ``pts3d_joints_views`` per track (see ``joints_views_list``), ``person_ids`` = the id
    repeated once per entry of the track's view dict, ``camera_ids`` / ``pts`` = cameras and 2-D poses matched
    this frame in the dict's insertion order.  Needs the ``nviews``, ``assoc`` and ``vlist`` outputs and the
    detection tensor ``dets (S,T,V,D,J,3)``; without ``vlist`` the view order falls back to camera order."""
    k = int(_np(out["count"])[s, t])
    ids = _np(out["ids"])[s, t, :k].astype(np.int64)
    # np.array(list of (3, J)) like the reference: shape (0,) on a frame without reported tracks
    pts3d = np.array([np.transpose(p) for p in _np(out["joints"])[s, t, :k].astype(np.float64)])
    vl = _np(out["vlist"])[s, t, :k] if out.get("vlist") is not None else None
    views = []
    if out.get("nviews") is not None:
        nv = _np(out["nviews"])[s, t, :k]
        for r, row in enumerate(nv):
            used = int(vl[r, 0]) if vl is not None else max(n_views, int(row.max(initial=0)))
            views.append(joints_views_list(row, used))
    camera_ids, pts, person_ids = [], [], []
    if dets is not None and out.get("assoc") is not None:
        assoc = _np(out["assoc"])[s, t]
        d = _np(dets)[s, t]
        for r, tid in enumerate(ids):
            if vl is not None:
                n_list = int(vl[r, 1])
                cams = [int(b & 0x7f) for b in vl[r, 2:2 + n_list] if b & 0x80]      # dict order, matched this frame
            else:
                n_list = None
                cams = [int(c) for c in range(assoc.shape[0]) if (assoc[c] == tid).any()]
            camera_ids.append(cams)
            pts.append([d[c, int(np.nonzero(assoc[c] == tid)[0][0])].astype(np.float64) for c in cams])
            person_ids.append([int(tid)] * (n_list if n_list is not None else len(cams)))
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
