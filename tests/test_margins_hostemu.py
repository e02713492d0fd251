"""Decision-margin counters (-DPAM_MARGIN) on the host build of the kernel source: every decision
class that occurs in a noisy stream leaves a finite, positive minimum; classes that never occur stay at +inf."""
import ctypes as C
import os
import subprocess

import numpy as np

from tests import util
from pam_b200 import camera, synth

_HERE = os.path.dirname(os.path.abspath(__file__))


def test_margins_are_recorded_per_sequence():
    so = os.path.join(_HERE, "hostemu", "_build", "libpam_hostemu_margin.so")
    src = os.path.join(_HERE, "hostemu", "hostemu.cpp")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-array-bounds",
                           "-DPAM_MARGIN", "-o", so, src])
    lib = C.CDLL(so)
    streams = [synth.make_stream("shelf", 60 + s, 80, miss_prob=0.1, outlier_prob=0.08, enter_stagger=10) for s in range(2)]
    cfg = util.stream_config(streams[0], max_tracks=8)
    L = util._capi.PamStateLayout()
    assert lib.hostemu_state_layout(C.byref(cfg), C.byref(L)) == 0
    cams = camera.GetCameraParameters(streams[0].rig)
    P, RK, pos, F = camera.pack_cameras(cams)
    dets = np.ascontiguousarray(np.stack([s.dets for s in streams]))
    counts = np.ascontiguousarray(np.stack([s.counts for s in streams]))
    S, T = dets.shape[:2]
    out = util.alloc_outputs(cfg, S, T)
    state = np.zeros(S * L.seq_bytes, np.uint8)
    status = np.zeros(S, np.int32)
    rc = lib.hostemu_track_sequences(C.byref(cfg), util.ptr(P), util.ptr(RK), util.ptr(pos), util.ptr(F), S, T, 0,
                                     util.ptr(dets), util.ptr(counts), util.ptr(out["count"]), util.ptr(out["ids"]),
                                     util.ptr(out["joints"]), util.ptr(out["nviews"]), util.ptr(out["assoc"]),
                                     util.ptr(status), util.ptr(state), None)
    assert rc == 0 and status.tolist() == [0, 0]
    m = np.zeros((S, L.n_margins))
    assert lib.hostemu_margins(C.byref(cfg), util.ptr(state), S, util.ptr(m)) == 0
    assert L.n_margins == 8
    # association c, update-mode A, ray rule, believe threshold, smallest positive affinity occur in every such stream
    for k in (0, 1, 2, 3, 7):
        assert np.all(np.isfinite(m[:, k])) and np.all(m[:, k] > 0), (k, m[:, k])
    assert np.all(m[:, 0] < 1.0) and np.all(m[:, 7] <= 1.0)
    # no decision of these streams sits within rounding distance of its threshold
    assert m[np.isfinite(m)].min() > 1e-9
