"""Batched, device-resident tracker: S independent sequences x T frames per call.

Thin host wrapper over the C ABI (``include/pam.h``): ``pam_track_sequences`` runs one CTA per
sequence over all T frames with the tracker state (tracks, view lists, smoothed-pose history,
velocities) resident in HBM; nothing returns to the host between frames.  torch is used only to own
device memory and streams."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _capi, camera as _camera

TENTATIVE, CONFIRMED, DELETED = 1, 2, 3
VLIST_BYTES = 10          # include/pam.h PAM_VLIST_BYTES


PAM_E_CAPACITY = -4
WARN_TRACKS, WARN_HYPOTHESES, WARN_DETECTIONS = 1, 2, 4


class PamError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libpam status {status}: {message}")
        self.status = status


class PamCapacityWarning(UserWarning):
    """A sequence needed more track slots / hypotheses / detections than configured: the excess was
    dropped for that frame and tracking went on (the reference has no such limits)."""


def _check(lib, handle, rc):
    if rc != 0:
        msg = lib.pam_last_error(handle)
        raise PamError(rc, (msg or b"").decode() or lib.pam_status_string(rc).decode())


def _np_ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class SequenceTracker:
    """``IterativeTracker`` semantics (src/tracking/IterativeTracker.py:34-180) for a batch of
    independent sequences that share one camera rig."""

    def __init__(self, cameras: Sequence, params, num_sequences: int = 1, max_detections: int = 8,
                 max_tracks: int = 8, arm_joints: Sequence[int] = (9, 10), min_valid_joints: int = 10,
                 device: int = 0, max_report: int = 0):
        """``max_report``: rows per frame of the output tensors (0 = ``max_tracks``); a frame that reports more tracks
        keeps the first ``max_report`` rows, ``count`` still tells the true number."""
        self.lib = _capi.load_library()
        self.cfg = _capi.make_config(params, len(cameras), max_detections, max_tracks, arm_joints, min_valid_joints,
                                     max_report=max_report)
        self.out_rows = max_report if 0 < max_report <= max_tracks else max_tracks
        self.S = int(num_sequences)
        self.device = int(device)
        self.handle = C.c_void_p()
        rc = self.lib.pam_create(C.byref(self.cfg), self.device, C.byref(self.handle))
        _check(self.lib, None, rc)
        self._cam_arrays = _camera.pack_cameras(cameras)
        _check(self.lib, self.handle, self.lib.pam_set_cameras(self.handle, *[_np_ptr(a) for a in self._cam_arrays]))
        self.layout = _capi.PamStateLayout()
        _check(self.lib, self.handle, self.lib.pam_get_state_layout(self.handle, C.byref(self.layout)))
        self.next_frame = 0
        self._state = None   # torch uint8 tensor on the device (device-pointer path)

    # -- life-cycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.pam_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.pam_launch_count(self.handle))

    def _torch(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("SequenceTracker needs a CUDA device (no CPU fallback)")
        return torch

    def restart(self):
        """``track_restart`` for every sequence."""
        torch = self._torch()
        if self._state is None:
            self._state = torch.empty(self.S * self.layout.seq_bytes, dtype=torch.uint8, device=f"cuda:{self.device}")
        st = torch.cuda.current_stream(self.device).cuda_stream
        _check(self.lib, self.handle, self.lib.pam_track_reset(self.handle, C.c_void_p(self._state.data_ptr()), self.S,
                                                               C.c_void_p(st)))
        self.next_frame = 0

    # -- device-pointer path ------------------------------------------------------------------
    def alloc_outputs(self, T: int, nviews: bool = True, assoc: bool = False, timing: bool = False, vlist: bool = False):
        torch = self._torch()
        dev = f"cuda:{self.device}"
        c = self.cfg
        out = dict(count=torch.empty((self.S, T), dtype=torch.int32, device=dev),
                   ids=torch.empty((self.S, T, self.out_rows), dtype=torch.int32, device=dev),
                   joints=torch.empty((self.S, T, self.out_rows, c.num_joints, 3), dtype=torch.float32, device=dev))
        out["nviews"] = torch.empty((self.S, T, self.out_rows, c.num_joints), dtype=torch.uint8, device=dev) if nviews else None
        out["assoc"] = torch.empty((self.S, T, c.num_cameras, c.max_detections), dtype=torch.int32, device=dev) if assoc else None
        if timing:
            out["timing"] = torch.zeros((self.S, T, 4), dtype=torch.int32, device=dev)
        if vlist:
            out["vlist"] = torch.zeros((self.S, T, self.out_rows, VLIST_BYTES), dtype=torch.uint8, device=dev)
        return out

    def run(self, dets, counts, out: Optional[dict] = None, frame0: Optional[int] = None, nviews=True, assoc=False,
            timing=False, vlist=False):
        """``dets`` (S,T,V,D,J,3) float32 and ``counts`` (S,T,V) int32 CUDA tensors.  Asynchronous on
        torch's current stream; returns the dict of output tensors."""
        torch = self._torch()
        c = self.cfg
        S, T = dets.shape[0], dets.shape[1]
        assert S == self.S and tuple(dets.shape[2:]) == (c.num_cameras, c.max_detections, c.num_joints, 3), dets.shape
        assert dets.dtype == torch.float32 and counts.dtype == torch.int32 and dets.is_contiguous() and counts.is_contiguous()
        assert dets.is_cuda and counts.is_cuda
        if self._state is None:
            self.restart()
        if out is None:
            out = self.alloc_outputs(T, nviews, assoc, timing, vlist)
        f0 = self.next_frame if frame0 is None else frame0
        st = torch.cuda.current_stream(self.device).cuda_stream
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        rc = self.lib.pam_track_sequences(self.handle, p(self._state), S, T, f0, p(dets), p(counts), p(out["count"]),
                                          p(out["ids"]), p(out["joints"]), p(out.get("nviews")), p(out.get("assoc")),
                                          p(out.get("timing")), p(out.get("vlist")), C.c_void_p(st))
        _check(self.lib, self.handle, rc)
        self.next_frame = f0 + T
        return out

    def check(self, strict: bool = True, host_path: bool = False):
        """Synchronise and return the per-sequence status words (hard error code | PAM_WARN bits << 8).
        A hard error always raises; a capacity warning (excess tracks / hypotheses / detections were dropped
        for a frame, tracking went on) raises when ``strict`` and is a ``PamCapacityWarning`` otherwise."""
        S = self._host_S if host_path else self.S
        status = np.zeros(S, np.int32)
        if host_path:
            rc = self.lib.pam_track_host_status(self.handle, S, _np_ptr(status))
        else:
            torch = self._torch()
            st = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.pam_track_status(self.handle, C.c_void_p(self._state.data_ptr()), S, _np_ptr(status), C.c_void_p(st))
        if rc == PAM_E_CAPACITY and not strict:
            import warnings
            warnings.warn((self.lib.pam_last_error(self.handle) or b"").decode(), PamCapacityWarning, stacklevel=2)
            return status
        _check(self.lib, self.handle, rc)
        return status

    def margins(self):
        """Decision margins per sequence (PAM_MARGIN builds of the library only): array (S, n_margins)."""
        torch = self._torch()
        out = np.zeros((self.S, self.layout.n_margins), np.float64)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _check(self.lib, self.handle, self.lib.pam_track_margins(self.handle, C.c_void_p(self._state.data_ptr()), self.S,
                                                                 _np_ptr(out), C.c_void_p(st)))
        return out

    def launch_info(self, S: Optional[int] = None) -> dict:
        """Launch shape the tracker uses for ``S`` sequences (defaults to this tracker's batch)."""
        v = np.zeros(8, np.int32)
        _check(self.lib, self.handle, self.lib.pam_track_launch_info(self.handle, int(self.S if S is None else S), _np_ptr(v)))
        keys = ("warps_per_sequence", "sequences_per_cta", "threads_per_cta", "ctas_per_sm", "smem_per_cta",
                "registers_per_thread", "arena_bytes_per_sequence", "detection_buffers")
        return dict(zip(keys, (int(x) for x in v)))

    def sm_clock_khz(self) -> int:
        v = C.c_int32(0)
        _check(self.lib, self.handle, self.lib.pam_sm_clock_khz(self.handle, C.byref(v)))
        return int(v.value)

    # -- host-buffer path (what a reference-side caller would use) ------------------------------
    def run_host(self, dets: np.ndarray, counts: np.ndarray, frame0: Optional[int] = None, fresh: bool = False,
                 nviews: bool = True, assoc: bool = False, out: Optional[dict] = None, timing: bool = False,
                 vlist: bool = False):
        """numpy in, numpy out through ``pam_track_sequences_host`` (H2D + kernel + D2H + sync)."""
        c = self.cfg
        dets = np.ascontiguousarray(dets, np.float32)
        counts = np.ascontiguousarray(counts, np.int32)
        S, T = dets.shape[0], dets.shape[1]
        assert tuple(dets.shape[2:]) == (c.num_cameras, c.max_detections, c.num_joints, 3), dets.shape
        if out is None:
            R = self.out_rows
            out = dict(count=np.empty((S, T), np.int32), ids=np.empty((S, T, R), np.int32),
                       joints=np.empty((S, T, R, c.num_joints, 3), np.float32),
                       nviews=np.empty((S, T, R, c.num_joints), np.uint8) if nviews else None,
                       assoc=np.empty((S, T, c.num_cameras, c.max_detections), np.int32) if assoc else None,
                       timing=np.zeros((S, T, 4), np.int32) if timing else None,
                       vlist=np.zeros((S, T, R, VLIST_BYTES), np.uint8) if vlist else None)
        if fresh:
            self.next_frame = 0
        f0 = self.next_frame if frame0 is None else frame0
        rc = self.lib.pam_track_sequences_host(self.handle, S, T, f0, 1 if fresh else 0, _np_ptr(dets), _np_ptr(counts),
                                               _np_ptr(out["count"]), _np_ptr(out["ids"]), _np_ptr(out["joints"]),
                                               _np_ptr(out.get("nviews")), _np_ptr(out.get("assoc")), _np_ptr(out.get("timing")),
                                               _np_ptr(out.get("vlist")))
        _check(self.lib, self.handle, rc)
        self.next_frame = f0 + T
        self._host_S = S
        return out

    # -- stream mode: one frame per call through the resident kernel -----------------------------
    def open_stream(self, fresh: bool = True) -> "FrameStream":
        """Start the resident per-frame kernel of this tracker's (single) sequence and return the zero-copy
        views of its command / result slot (``include/pam.h``: pam_stream_open)."""
        assert self.S == 1, "stream mode tracks one sequence"
        _check(self.lib, self.handle, self.lib.pam_stream_open(self.handle, 1 if fresh else 0))
        self._host_S = 1
        if fresh:
            self.next_frame = 0
        return FrameStream(self)

    # -- state read-back (IterTrack read surface) ---------------------------------------------
    def read_state(self, host_path: bool = False):
        """Parse the tracker state of every sequence: list (per sequence) of track dicts in
        track-list order with the reference's IterTrack fields."""
        L = self.layout
        if host_path:
            S = self._host_S
            blob = np.empty(S * L.seq_bytes, np.uint8)
            _check(self.lib, self.handle, self.lib.pam_track_state_to_host(self.handle, S, _np_ptr(blob)))
        else:
            S = self.S
            self._torch().cuda.synchronize(self.device)
            blob = self._state.cpu().numpy()
        return parse_state(blob, S, L, self.cfg)


class FrameStream:
    """numpy views over the pinned, device-mapped slot of the resident kernel.  Write the frame's detections PACKED
    (cameras back to back) into ``packed (V*D, J, 3)`` -- ``set_frame()`` does it from the reference's per-camera
    list -- and the per-camera ``counts (V,)``, call ``step(frame_id)`` (or ``submit`` / ``wait``), then read
    ``count[0]``, ``ids``, ``joints``, ``nviews``, ``assoc``, ``timing`` (SM cycles: association, update,
    initialisation, total)."""

    def __init__(self, trk: "SequenceTracker"):
        self.trk = trk
        v = _capi.PamStreamViews()
        _check(trk.lib, trk.handle, trk.lib.pam_stream_buffers(trk.handle, C.byref(v)))
        c = trk.cfg
        V, D, J, MT = c.num_cameras, c.max_detections, c.num_joints, trk.out_rows
        self.V, self.D = V, D

        def view(ptr, ctype, shape):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=shape)

        self.packed = view(v.dets, C.c_float, (V * D, J, 3))
        self.counts = view(v.counts, C.c_int32, (V,))
        self.count = view(v.out_count, C.c_int32, (1,))
        self.ids = view(v.out_ids, C.c_int32, (MT,))
        self.joints = view(v.out_joints, C.c_float, (MT, J, 3))
        self.nviews = view(v.out_nviews, C.c_uint8, (MT, J))
        self.assoc = view(v.out_assoc, C.c_int32, (V, D))
        self.timing = view(v.out_timing, C.c_int32, (8,))   # [4:7]: protocol timing of the resident kernel
        self.status = view(v.out_status, C.c_int32, (1,))
        self.vlist = view(v.out_vlist, C.c_uint8, (MT, VLIST_BYTES))
        self._lib, self._handle = trk.lib, trk.handle
        self._submit, self._wait = trk.lib.pam_stream_submit, trk.lib.pam_stream_wait

    def set_frame(self, detections_list):
        """Per camera ``(m, J, 3)`` arrays (any float dtype) -> the slot.  Returns the number of rows written."""
        D = self.D
        lst = [d[:D] for d in detections_list if len(d)]          # at most D detections per camera reach the device
        self.counts[:] = [min(len(d), D) for d in detections_list]
        n = sum(len(d) for d in lst)
        if n:
            np.concatenate(lst, axis=0, out=self.packed[:n], casting="same_kind")
        return n

    def set_padded(self, dets, counts):
        """``dets (V, D, J, 3)`` zero padded + ``counts (V,)`` (one frame of the batched layout) -> the slot."""
        self.set_frame([dets[c, :counts[c]] for c in range(self.V)])

    def submit(self, frame_id: int):
        rc = self._submit(self._handle, frame_id)
        if rc != 0:
            _check(self._lib, self._handle, rc)

    def wait(self):
        rc = self._wait(self._handle)
        if rc != 0:
            _check(self._lib, self._handle, rc)

    def step(self, frame_id: int):
        self.submit(frame_id)
        self.wait()

    def close(self):
        _check(self._lib, self._handle, self._lib.pam_stream_close(self._handle))


def parse_state(blob: np.ndarray, S: int, L, cfg):
    J, V, MT, H = cfg.num_joints, cfg.num_cameras, cfg.max_tracks, L.hist_ring
    seqs = []
    for s in range(S):
        base = blob[s * L.seq_bytes:(s + 1) * L.seq_bytes]
        hdr = base[L.off_header:L.off_header + 4 * L.header_ints].view(np.int32)
        ntracks, next_id, status, warn = int(hdr[0]), int(hdr[1]), int(hdr[2]), int(hdr[5])
        order = base[L.off_header + 4 * L.header_ints:L.off_header + 4 * L.header_ints + ntracks].view(np.int8)
        meta = base[L.off_meta:L.off_meta + 4 * L.meta_ints * MT].view(np.int32).reshape(MT, L.meta_ints)
        hist = base[L.off_hist:L.off_hist + 8 * MT * H * J * 3].view(np.float64).reshape(MT, H, J, 3)
        view = base[L.off_view:L.off_view + 4 * MT * V * J * 3].view(np.float32).reshape(MT, V, J, 3)
        vel = base[L.off_vel:L.off_vel + 4 * MT * J * 3].view(np.float32).reshape(MT, J, 3)
        nv = base[L.off_nviews:L.off_nviews + MT * J].reshape(MT, J)
        tracks = []
        for slot in order:
            m = meta[slot]
            nviews, hs, hl, vt_last = int(m[6]), int(m[7]), int(m[8]), int(m[9])
            vc = m[10:10 + L.max_views]
            vt = m[10 + L.max_views:10 + 2 * L.max_views]
            ht = m[10 + 2 * L.max_views:10 + 2 * L.max_views + H]
            poses2d = {int(vc[k]): dict(time=int(vt[k]), pose=view[slot, k].astype(np.float64)) for k in range(nviews)}
            ring = [(hs + i) % H for i in range(hl)]
            poses3d = [dict(time=int(ht[r]), pose3d=hist[slot, r].copy()) for r in ring]
            tracks.append(dict(track_id=int(m[0]), hits=int(m[1]), age=int(m[2]), time_since_update=int(m[3]),
                               state=int(m[4]), already_update=bool(m[5]), poses2d=poses2d, poses3d=poses3d,
                               velocity_3d=vel[slot].copy(), joint_views=nv[slot].copy(), views_used=vt_last))
        seqs.append(dict(tracks=tracks, next_id=next_id, status=status, warn=warn, warn_frames=int(hdr[6])))
    return seqs
