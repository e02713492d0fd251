"""CPU check of the KERNEL SOURCE: csrc/pam_track.h compiled for the host with one "thread"
(tests/hostemu, test infrastructure only) must reproduce the oracle's decisions.  This is how the
algorithmic logic of the CUDA path is iterated on in the GPU-less build container; the real parity
tests are the -m gpu ones."""
import pytest

from tests import util
from pam_b200 import synth

CASES = [
    ("shelf", {}, 120),
    ("shelf", dict(enter_stagger=25, miss_prob=0.1, outlier_prob=0.05, absences=[(1, 60, 90)]), 200),
    ("campus", dict(miss_prob=0.05, outlier_prob=0.03), 150),
    ("shelf17", dict(miss_prob=0.03, outlier_prob=0.02), 80),
    ("panoptic", dict(miss_prob=0.05, outlier_prob=0.03), 60),
]


@pytest.mark.parametrize("shape,kw,T", CASES)
def test_kernel_source_on_host_matches_oracle(shape, kw, T):
    st = synth.make_stream(shape, 7, T, **kw)
    cfg = util.stream_config(st, max_tracks=12)
    out = util.run_hostemu([st], cfg)
    assert out["status"].tolist() == [0]
    oo, oa, _ = util.run_oracle(st)
    worst = util.compare_with_oracle(out, 0, st, oo, oa)
    assert out["count"].sum() > 0 and worst < 5e-4


@pytest.mark.parametrize("shape,mult", [("shelf", 6.0), ("campus", 6.0), ("panoptic", 4.0)])
@pytest.mark.parametrize("twopass", ["0", "1"])
def test_wide_association_threshold_exercises_the_assignment_solver(shape, mult, twopass, monkeypatch):
    """alpha2d several times the dataset value: most tracks see more than one candidate detection per camera, so a
    large share of the frames goes through the full assignment solver.  The kernel only evaluates the SIGN of an
    affinity until a camera needs the solver (then the values of its positive entries are filled in): decisions must
    stay those of the reference, which computes every value (IterativeTracker.py:139-160)."""
    monkeypatch.setenv("PAM_HOSTEMU_TWOPASS", twopass)
    st = synth.make_stream(shape, 5, 160, miss_prob=0.1, outlier_prob=0.1)
    p = synth.tracker_params(shape)
    p["alpha2d"] = p["alpha2d"] * mult
    cfg = util.stream_config(st, max_tracks=16 if shape == "panoptic" else 8, params=p)
    out = util.run_hostemu([st], cfg)
    assert out["status"].tolist() == [0]
    oo, oa, _ = util.run_oracle(st, params=p)
    assert util.compare_with_oracle(out, 0, st, oo, oa) < 5e-4


def test_guard_band_fallback_of_the_sign_only_affinity():
    """Host build with the guard band of the sign-only association test widened from 2^-24 to 0.6 (-DPAM_SIGN_BAND):
    every joint between 0.63 and 1.26 of the threshold is then decided by the exact out-of-line expression.  Same
    decisions as the oracle, i.e. the fallback agrees with the fast comparison wherever both apply."""
    import ctypes as C
    import os
    import subprocess
    import numpy as np
    from pam_b200 import camera
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostemu", "_build", "libpam_hostemu_band.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-array-bounds",
                           "-DPAM_SIGN_BAND=0.6", "-o", so, os.path.join(here, "hostemu", "hostemu.cpp")])
    lib = C.CDLL(so)
    st = synth.make_stream("shelf", 11, 150, miss_prob=0.1, outlier_prob=0.1)
    p = synth.tracker_params("shelf")
    p["alpha2d"] = 12.0                                    # detections scatter around the threshold
    cfg = util.stream_config(st, max_tracks=32, params=p)          # short-lived tracks come and go: 8 slots overflow
    P, RK, pos, F = camera.pack_cameras(camera.GetCameraParameters(st.rig))
    dets, counts = np.ascontiguousarray(st.dets[None]), np.ascontiguousarray(st.counts[None])
    out = util.alloc_outputs(cfg, 1, st.T)
    status = np.zeros(1, np.int32)
    rc = lib.hostemu_track_sequences(C.byref(cfg), util.ptr(P), util.ptr(RK), util.ptr(pos), util.ptr(F), 1, st.T, 0,
                                     util.ptr(dets), util.ptr(counts), util.ptr(out["count"]), util.ptr(out["ids"]),
                                     util.ptr(out["joints"]), util.ptr(out["nviews"]), util.ptr(out["assoc"]),
                                     util.ptr(status), None, util.ptr(out["vlist"]))
    assert rc == 0 and status[0] == 0
    oo, oa, _ = util.run_oracle(st, params=p)
    assert util.compare_with_oracle(out, 0, st, oo, oa) < 5e-4 and out["count"].sum() > 0


def test_kernel_source_chunked_launches_carry_state():
    """Frames fed in several 'launches' (tracker state carried in the state buffer, stale views first read
    from the launch's own input, then from the persisted copies) give the result of a single launch."""
    import ctypes as C
    import numpy as np
    from tests.hostemu import build as hb
    from pam_b200 import _capi, camera
    st = synth.make_stream("shelf", 31, 90, miss_prob=0.25, outlier_prob=0.05)
    cfg = util.stream_config(st, max_tracks=12)
    one = util.run_hostemu([st], cfg)
    lib = hb.load()
    L = _capi.PamStateLayout()
    assert lib.hostemu_state_layout(C.byref(cfg), C.byref(L)) == 0
    state = np.zeros(L.seq_bytes, np.uint8)
    P, RK, pos, F = camera.pack_cameras(camera.GetCameraParameters(st.rig))
    parts = []
    for a, b in ((0, 1), (1, 2), (2, 17), (17, 18), (18, 60), (60, 90)):
        dets = np.ascontiguousarray(st.dets[None, a:b]); counts = np.ascontiguousarray(st.counts[None, a:b])
        out = util.alloc_outputs(cfg, 1, b - a)
        status = np.zeros(1, np.int32)
        rc = lib.hostemu_track_sequences(C.byref(cfg), util.ptr(P), util.ptr(RK), util.ptr(pos), util.ptr(F), 1, b - a, a,
                                         util.ptr(dets), util.ptr(counts), util.ptr(out["count"]), util.ptr(out["ids"]),
                                         util.ptr(out["joints"]), util.ptr(out["nviews"]), util.ptr(out["assoc"]),
                                         util.ptr(status), util.ptr(state), util.ptr(out["vlist"]))
        assert rc == 0 and status[0] == 0
        parts.append(out)
    for k in ("count", "assoc"):
        assert np.array_equal(np.concatenate([p[k] for p in parts], 1), one[k]), k
    cnt = one["count"]
    for k in ("ids", "joints", "nviews"):
        cat = np.concatenate([p[k] for p in parts], 1)
        for t in range(st.T):
            assert np.array_equal(cat[0, t, :cnt[0, t]], one[k][0, t, :cnt[0, t]]), (k, t)
