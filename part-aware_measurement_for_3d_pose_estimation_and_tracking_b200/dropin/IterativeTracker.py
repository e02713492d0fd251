"""GPU-backed ``IterativeTracker`` with the reference's call surface
(src/tracking/IterativeTracker.py:21-50, 115-180, 182-287).

``tracking()`` hands ONE frame of ONE sequence to the resident kernel of the library's stream mode
(``pam_stream_step``: the detections are written into a pinned, device-mapped slot, the kernel keeps the
tracker state on chip and writes the reported tracks back; no launch, no copies, no state round trip per
call).  ``.tracks`` reads the state back on demand (this parks the resident kernel; the next ``tracking()``
call restarts it) and exposes the IterTrack read surface."""
import time as _time

import numpy as np

from _pkg import camera as _camera, tracker as _tracker, capi as _capi, results as _results
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
        # joints_views of the newest pose as get_3dpose / get_3dpose_jf build it: one bucket per usable view of that
        # update (src/tracking/IterativeTracker.py:350-357); older history entries do not keep theirs on the device
        jv = _results.joints_views_list(d["joint_views"], d["views_used"])
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
    #: capacities of the device-resident state (the reference has none): track slots per sequence (<= 32) and
    #: detections per camera and frame (<= 16).  A frame that needs more DROPS the excess (the new track is not
    #: created / the detection is ignored), tracking goes on and a PamCapacityWarning is issued once.
    MAX_TRACKS = 12
    MAX_DETECTIONS = 8
    #: The device path ingests detections as float32 (the reference's come from a float32 pose net and are
    #: float32-exact).  True: a float64 value that float32 cannot represent raises ValueError instead of being
    #: rounded silently (parity with the reference is only claimed for representable input).
    STRICT_FLOAT32 = True

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
        self._stream = None
        self._cam_list_ref = None
        self._fresh = True
        self._warned = False

    def track_restart(self):
        self._pending = None
        self._unmatched = dict()
        self._ids_seen = set()
        self._tracks_cache = None
        self._fresh = True

    def _ensure(self, camera_list):
        if camera_list is self._cam_list_ref and not self._fresh and self._stream is not None:
            return                                     # same list object as last frame: nothing to set up
        if self._trk is None or self._cams is None or len(camera_list) != len(self._cams) or \
                any(a is not b for a, b in zip(camera_list, self._cams)):
            if self._trk is not None:
                self._trk.close()
            self._cams = list(camera_list)
            self._trk = _tracker.SequenceTracker(self._cams, self.args, 1, self.MAX_DETECTIONS, self.MAX_TRACKS,
                                                 self.ARM_JOINTS, self.MIN_VALID_JOINTS)
            self._fresh = True
            self._stream = None
            self._cycle_s = 1.0 / (1e3 * max(1, self._trk.sm_clock_khz()))
        self._cam_list_ref = camera_list
        if self._fresh or self._stream is None:
            self._stream = self._trk.open_stream(fresh=self._fresh)      # resident kernel, state on chip
            st = self._stream
            self._fast = _capi.load_pyfast()
            self._handle_addr = self._trk.handle.value
            self._packed_addr, self._counts_addr = st.packed.ctypes.data, st.counts.ctypes.data
            self._row = st.packed.shape[1] * 3
            self._fresh = False

    def tracking(self, frame_id, camera_list, frame_list, boxes_list, detections_list, build3D='TopDown'):
        """One frame (src/tracking/IterativeTracker.py:115-180).  Returns ``(asso_time, update_time, init_time)`` in
        seconds, measured on the device (SM cycle counters around the same three blocks the reference times)."""
        assert build3D == 'SVD', "Please modify BUILD3D to SVD when PERSON_MATCHER == Iterative"
        self.frame_list = frame_list
        self.build3D = build3D
        self.cam_num = len(camera_list)
        self._ensure(camera_list)
        st = self._stream
        self._tracks_cache = None
        self._pending = (frame_id, camera_list, boxes_list, detections_list)
        fast = self._fast
        if fast is not None:
            # one call: float64 -> float32 packing into the mapped slot (with the representability check), submit, wait
            try:
                rc, flags = fast.track_frame(self._handle_addr, int(frame_id), detections_list, self._packed_addr,
                                             self._counts_addr, st.V, st.D, self._row, 1 if self.STRICT_FLOAT32 else 0)
            except TypeError:
                fast = None                # python lists instead of arrays etc.: numpy path below
            else:
                if rc != 0:
                    _tracker._check(self._trk.lib, self._trk.handle, rc)
                if flags & 2 and self.STRICT_FLOAT32:
                    raise ValueError("detections are not float32-representable; the device path would round them "
                                     "(set IterativeTracker.STRICT_FLOAT32 = False to accept the rounding)")
                if flags & 1:
                    self._warn(f"more than MAX_DETECTIONS={self.MAX_DETECTIONS} detections in a camera: the first "
                               f"{self.MAX_DETECTIONS} were used")
        if fast is None:
            maxd = self.MAX_DETECTIONS
            detections_list = [np.asarray(d) for d in detections_list]
            if any(len(d) > maxd for d in detections_list):
                self._warn(f"more than MAX_DETECTIONS={maxd} detections in a camera: the first {maxd} are used")
                detections_list = [d[:maxd] for d in detections_list]
            lst = [d for d in detections_list if len(d)]
            st.counts[:] = [len(d) for d in detections_list]
            n = sum(len(d) for d in lst)
            if n:
                np.concatenate(lst, axis=0, out=st.packed[:n], casting="same_kind")
                if self.STRICT_FLOAT32 and not np.array_equal(st.packed[:n], np.concatenate(lst, axis=0)):
                    raise ValueError("detections are not float32-representable; the device path would round them "
                                     "(set IterativeTracker.STRICT_FLOAT32 = False to accept the rounding)")
            st.step(int(frame_id))
        if st.status[0] and not self._warned:
            self._warn("a frame needed more track slots / hypotheses / detections than configured "
                       f"(MAX_TRACKS={self.MAX_TRACKS}, MAX_DETECTIONS={self.MAX_DETECTIONS}); the excess was dropped")
        t_asso, t_update, t_init = st.timing[:3].tolist()
        cs = self._cycle_s
        return t_asso * cs, t_update * cs, t_init * cs

    # results of the latest frame, read from the mapped slot on demand (valid until the next tracking() call)
    @property
    def last_ids(self):
        return self._stream.ids[:int(self._stream.count[0])].copy()

    @property
    def last_joints(self):
        return self._stream.joints[:int(self._stream.count[0])].astype(np.float64)

    @property
    def last_nviews(self):
        return self._stream.nviews[:int(self._stream.count[0])].copy()

    def _warn(self, msg):
        if not self._warned:
            import warnings
            warnings.warn(msg, _tracker.PamCapacityWarning, stacklevel=3)
            self._warned = True

    @property
    def unmatched(self):
        """Leftover detections per camera, as the reference leaves them after ``init_target_GD``
        (src/tracking/IterativeTracker.py:56-61, 163-167); built on first access after a frame."""
        if self._pending is not None:
            frame_id, camera_list, boxes_list, detections_list = self._pending
            assoc = self._stream.assoc.copy()
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
