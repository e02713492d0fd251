"""Shared helpers for the parity tests."""
from __future__ import annotations

import ctypes as C

import numpy as np

import pam_b200  # noqa: F401  (alias for the hyphenated package directory)
from pam_b200 import _capi, camera, synth


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def stream_config(stream, max_tracks=8, max_detections=None, params=None, min_valid_joints=10):
    sh = stream.shape
    params = params or synth.tracker_params(sh)
    D = stream.dets.shape[2] if max_detections is None else max_detections
    return _capi.make_config(params, sh.V, D, max_tracks, arm_joints=sh.arm_joints,
                             min_valid_joints=min_valid_joints)


def alloc_outputs(cfg, S, T, want_assoc=True):
    MT, J, V, D = cfg.max_tracks, cfg.num_joints, cfg.num_cameras, cfg.max_detections
    return dict(count=np.zeros((S, T), np.int32), ids=np.full((S, T, MT), -1, np.int32),
                joints=np.zeros((S, T, MT, J, 3), np.float32), nviews=np.zeros((S, T, MT, J), np.uint8),
                assoc=np.full((S, T, V, D), -1, np.int32) if want_assoc else None,
                vlist=np.zeros((S, T, MT, 10), np.uint8))


def run_hostemu(streams, cfg, frame0=0):
    """Run the host-compiled kernel source over a list of streams that share one rig."""
    from tests.hostemu import build as hb
    lib = hb.load()
    cams = camera.GetCameraParameters(streams[0].rig)
    P, RK, pos, F = camera.pack_cameras(cams)
    dets = np.ascontiguousarray(np.stack([s.dets for s in streams]))
    counts = np.ascontiguousarray(np.stack([s.counts for s in streams]))
    S, T = dets.shape[0], dets.shape[1]
    out = alloc_outputs(cfg, S, T)
    status = np.zeros(S, np.int32)
    rc = lib.hostemu_track_sequences(C.byref(cfg), ptr(P), ptr(RK), ptr(pos), ptr(F), S, T, frame0, ptr(dets),
                                     ptr(counts), ptr(out["count"]), ptr(out["ids"]), ptr(out["joints"]),
                                     ptr(out["nviews"]), ptr(out["assoc"]), ptr(status), None, ptr(out["vlist"]))
    assert rc == 0, rc
    out["status"] = status
    return out


def run_oracle(stream, params=None, min_valid_joints=10, T=None):
    from oracle import generic
    params = params or synth.tracker_params(stream.shape)
    return generic.run_stream(stream, params, stream.shape.arm_joints, min_valid_joints, T=T, trace=True)


def compare_with_oracle(out, s, stream, oracle_out, oracle_assoc, tol_abs=5e-4, tol_rel=1e-3):
    """Track ids, reported-track counts, per-joint view counts and association decisions must be
    identical; 3-D joints within 0.5 mm absolute or 1e-3 relative (BASELINE.json north_star).
    Returns the max absolute joint deviation in metres."""
    T = len(oracle_out)
    worst = 0.0
    for t in range(T):
        ids, joints, views = oracle_out[t]
        k = int(out["count"][s, t])
        assert k == len(ids), f"frame {t}: {k} tracks reported, oracle {len(ids)}"
        assert np.array_equal(out["ids"][s, t, :k], ids), f"frame {t}: ids {out['ids'][s, t, :k]} vs {ids}"
        if k:
            assert np.array_equal(out["nviews"][s, t, :k], views), f"frame {t}: per-joint view counts differ"
            got = out["joints"][s, t, :k].astype(np.float64)
            err = np.abs(got - joints)
            lim = np.maximum(tol_abs, tol_rel * np.abs(joints))
            assert np.all(err <= lim), f"frame {t}: joint error {err.max():.3e} m"
            worst = max(worst, float(err.max()))
        if out.get("assoc") is not None:
            for c, a in enumerate(oracle_assoc[t]):
                assert np.array_equal(out["assoc"][s, t, c, :len(a)], a), f"frame {t} cam {c}: association differs"
    return worst


# ------------------------------------------------------------------------------------------------
# golden vectors + drop-in module loading
# ------------------------------------------------------------------------------------------------
import os as _os
import sys as _sys

GOLDEN = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")


def golden_streams():
    z = np.load(_os.path.join(GOLDEN, "streams.npz"))
    out = []
    for n in range(int(z["n_streams"])):
        p = f"s{n}_"
        shape = synth.SHAPES[str(z[p + "shape"])]
        rig = dict(P=z[p + "P"], K=z[p + "K"], RT=z[p + "RT"], width=shape.width, height=shape.height)
        st = synth.Stream(shape, -1, rig, z[p + "dets"], z[p + "counts"], None, None)
        out.append((st, {k: z[p + k] for k in ("count", "ids", "joints", "views", "assoc")}))
    return out


def golden_as_oracle_lists(g):
    """Golden arrays -> the per-frame lists compare_with_oracle() expects."""
    T = len(g["count"])
    frames = [(g["ids"][t, :g["count"][t]], g["joints"][t, :g["count"][t]], g["views"][t, :g["count"][t]]) for t in range(T)]
    assoc = [[g["assoc"][t, c] for c in range(g["assoc"].shape[1])] for t in range(T)]
    return frames, assoc


def golden_results():
    """The reference facade's per-frame return tuple for the first golden stream (tests/golden/make_golden.py)."""
    import gzip
    import pickle
    with gzip.open(_os.path.join(GOLDEN, "results.pkl.gz"), "rb") as f:
        return pickle.load(f)


def check_results_against_facade(out, dets, g):
    """results.person_track_output on tracker outputs (count, ids, joints, nviews, assoc, vlist) against the tuple
    PersonTrack_Project3DPose returned for the same stream (src/ivclabpose.py:259-287)."""
    from pam_b200 import results
    for t, fr in enumerate(g["frames"]):
        cam_ids, pts, person_ids, pts3d, jviews, p3ids = results.person_track_output(out, 0, t, dets=dets)
        assert [int(x) for x in p3ids] == fr["person3d_ids"], t
        assert [list(x) for x in cam_ids] == fr["camera_ids"], (t, [list(x) for x in cam_ids], fr["camera_ids"])
        assert person_ids == fr["person_ids"], t
        assert jviews == fr["joints_views"], (t, jviews, fr["joints_views"])
        assert pts3d.shape == fr["pts3d"].shape and (pts3d.size == 0 or np.abs(pts3d - fr["pts3d"]).max() < 5e-4), t
        for a, b in zip(pts, fr["pts"]):
            assert len(a) == len(b) and all(np.array_equal(np.asarray(x, np.float32), y) for x, y in zip(a, b)), t


def golden_functions():
    return np.load(_os.path.join(GOLDEN, "functions.npz"))


def load_dropin():
    """Import the flat drop-in modules (calculate, matching, construction, hypothesis,
    IterativeTracker) without leaving them in sys.modules under those generic names (pytest's
    own `hypothesis` plugin uses one of them)."""
    import importlib.util
    import types
    from pam_b200 import dropin
    here = _os.path.dirname(_os.path.abspath(dropin.__file__))   # not dropin.install(): keep sys.path clean
    names = ["_pkg", "calculate", "matching", "construction", "hypothesis", "IterativeTracker", "OneEuroFilter", "KalmanFilter"]
    saved = {k: _sys.modules.get(k) for k in names}
    ns = types.SimpleNamespace()
    try:
        for k in names:
            _sys.modules.pop(k, None)
        for k in names:
            spec = importlib.util.spec_from_file_location(k, _os.path.join(here, k + ".py"))
            mod = importlib.util.module_from_spec(spec)
            _sys.modules[k] = mod
            spec.loader.exec_module(mod)
            setattr(ns, k, mod)
    finally:
        for k in names:
            _sys.modules.pop(k, None)
            if saved[k] is not None:
                _sys.modules[k] = saved[k]
    return ns
