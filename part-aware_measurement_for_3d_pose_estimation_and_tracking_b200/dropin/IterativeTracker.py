"""GPU-backed ``IterativeTracker`` with the reference's call surface
(src/tracking/IterativeTracker.py:21-50, 115-180, 182-287).

``tracking()`` runs ONE frame of ONE sequence through ``pam_track_sequences_host`` (H2D of the
frame's detections, one kernel launch, D2H of the reported tracks); the tracker state stays in HBM
between calls.  ``.tracks`` reads the state back on demand and exposes the IterTrack read surface."""
import time as _time

import numpy as np

from _pkg import camera as _camera, tracker as _tracker
from calculate import get_believe


class TrackState:
    Tentative = 1
    Confirmed = 2
    Deleted = 3


class IterTrack:
    """Read-only view of one device-resident track (fields of src/tracking/IterativeTracker.py:194-204)."""

    def __init__(self, d, cameras):
        self.track_id = d["track_id"]
        self.hits, self.age, self.time_since_update = d["hits"], d["age"], d["time_since_update"]
        self.state, self.already_update = d["state"], d["already_update"]
        self.velocity_3d = d["velocity_3d"]
        self.joints = d["velocity_3d"].shape[0]
        self.poses2d = {cid: {"time": v["time"], "camera": cameras[cid], "pose": v["pose"]} for cid, v in d["poses2d"].items()}
        nv = d["joint_views"]
        jv = [[] for _ in range(max(len(cameras), 1))]
        for j, k in enumerate(nv):
            if k > 0:
                jv[int(k) - 1].append(j)
        self.poses3d = [{"time": p["time"], "pose3d": p["pose3d"], "joints_views": None} for p in d["poses3d"]]
        if self.poses3d:
            self.poses3d[-1]["joints_views"] = jv

    def is_tentative(self):
        return self.state == TrackState.Tentative

    def is_confirmed(self):
        return self.state == TrackState.Confirmed

    def is_deleted(self):
        return self.state == TrackState.Deleted


class IterativeTracker(object):
    #: joints smoothed with ARM_SIGMA and the valid-joint threshold the reference hard-wires for
    #: COCO-17 (src/tracking/IterativeTracker.py:382, :145); overridable for other skeletons
    ARM_JOINTS = (9, 10)
    MIN_VALID_JOINTS = 10
    MAX_TRACKS = 16
    MAX_DETECTIONS = 16

    def __init__(self, args):
        self.args = args
        self.cam_num = 0
        self.conf_threshold = self.args.conf_threshold
        self.num_joints = self.args.num_joints
        self.epi_threshold = self.args.epi_threshold
        self._pending = None
        self._unmatched = dict()
        self._ids_seen = set()
        self.build3D = None
        self._trk = None
        self._cams = None
        self._tracks_cache = None

    def track_restart(self):
        self._pending = None
        self._unmatched = dict()
        self._ids_seen = set()
        self._tracks_cache = None
        self._fresh = True

    def _ensure(self, camera_list):
        if self._trk is None or self._cams is None or len(camera_list) != len(self._cams) or \
                any(a is not b for a, b in zip(camera_list, self._cams)):
            if self._trk is not None:
                self._trk.close()
            self._cams = list(camera_list)
            self._trk = _tracker.SequenceTracker(self._cams, self.args, 1, self.MAX_DETECTIONS, self.MAX_TRACKS,
                                                 self.ARM_JOINTS, self.MIN_VALID_JOINTS)
            self._fresh = True
            V, D, J, MT = len(self._cams), self.MAX_DETECTIONS, self.num_joints, self.MAX_TRACKS
            self._dets = np.zeros((1, 1, V, D, J, 3), np.float32)
            self._counts = np.zeros((1, 1, V), np.int32)
            self._out = dict(count=np.zeros((1, 1), np.int32), ids=np.zeros((1, 1, MT), np.int32),
                             joints=np.zeros((1, 1, MT, J, 3), np.float32), nviews=np.zeros((1, 1, MT, J), np.uint8),
                             assoc=np.zeros((1, 1, V, D), np.int32))

    def tracking(self, frame_id, camera_list, frame_list, boxes_list, detections_list, build3D='TopDown'):
        assert build3D == 'SVD', "Please modify BUILD3D to SVD when PERSON_MATCHER == Iterative"
        self.frame_list = frame_list
        self.build3D = build3D
        self.cam_num = len(camera_list)
        self._ensure(camera_list)
        t0 = _time.time()
        self._counts[:] = 0
        for c, dets in enumerate(detections_list):
            m = len(dets)
            if m > self.MAX_DETECTIONS:
                raise ValueError(f"camera {c}: {m} detections exceed MAX_DETECTIONS={self.MAX_DETECTIONS}")
            if m:
                self._dets[0, 0, c, :m] = np.asarray(dets, dtype=np.float32)
            self._counts[0, 0, c] = m
        self._trk.run_host(self._dets, self._counts, frame0=int(frame_id), fresh=self._fresh, nviews=True, assoc=True,
                           out=self._out)
        self._fresh = False
        self._tracks_cache = None
        k = int(self._out["count"][0, 0])
        self.last_ids = self._out["ids"][0, 0, :k].copy()
        self.last_joints = self._out["joints"][0, 0, :k].astype(np.float64)
        self._pending = (frame_id, camera_list, boxes_list, detections_list, self._out["assoc"][0, 0].copy())
        return _time.time() - t0, 0.0, 0.0

    @property
    def unmatched(self):
        """Leftover detections per camera, as the reference leaves them after ``init_target_GD``
        (src/tracking/IterativeTracker.py:56-61, 163-167); built on first access after a frame."""
        if self._pending is not None:
            frame_id, camera_list, boxes_list, detections_list, assoc = self._pending
            self._pending = None
            for c, (camera, boxes, dets) in enumerate(zip(camera_list, boxes_list, detections_list)):
                m = len(dets)
                free = [d for d in range(m) if assoc[c, d] < 0]
                kept = [np.asarray(dets[d]) for d in free]
                if len(camera_list) >= 2:
                    kept = [d for d in kept if get_believe(d) > self.conf_threshold]
                bx = np.asarray(boxes)[free] if len(np.asarray(boxes)) == m and m else np.asarray(boxes)
                self._unmatched[camera.cid] = {'camera': camera, 'time': frame_id, 'bboxes': bx, 'detections': np.array(kept)}
        return self._unmatched

    @unmatched.setter
    def unmatched(self, value):
        self._unmatched = value
        self._pending = None

    @property
    def tracks(self):
        if self._trk is None:
            return []
        if self._tracks_cache is None:
            st = self._trk.read_state(host_path=True)[0]
            self._tracks_cache = [IterTrack(d, self._cams) for d in st["tracks"]]
            self._next_id = st["next_id"]
        return self._tracks_cache

    @property
    def tracks_ids(self):
        """Every id handed out so far (the reference never prunes this set, :113)."""
        if self._trk is None:
            return set()
        self.tracks  # refresh the state read-back
        return set(range(getattr(self, "_next_id", 0)))
