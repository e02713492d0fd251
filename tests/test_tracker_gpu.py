"""Parity of the persistent tracker kernel (pam_track_sequences, through the C ABI) with the CPU
oracle on seeded synthetic streams: identical track ids, reported-track sets, per-joint view
counts and association decisions; 3-D joints within 0.5 mm / 1e-3 relative."""
import numpy as np
import pytest

from tests import util
from pam_b200 import camera, synth, tracker

pytestmark = pytest.mark.gpu

CASES = [
    ("shelf", {}, 200),
    ("shelf", dict(enter_stagger=25, miss_prob=0.1, outlier_prob=0.05, absences=[(1, 60, 90)]), 300),
    ("campus", dict(miss_prob=0.05, outlier_prob=0.03), 300),
    ("shelf17", dict(miss_prob=0.03, outlier_prob=0.02), 150),
    ("panoptic", dict(miss_prob=0.05, outlier_prob=0.03), 150),
]


def _tracker_for(streams, max_tracks=12):
    st = streams[0]
    cams = camera.GetCameraParameters(st.rig)
    return tracker.SequenceTracker(cams, synth.tracker_params(st.shape), num_sequences=len(streams),
                                   max_detections=st.dets.shape[2], max_tracks=max_tracks,
                                   arm_joints=st.shape.arm_joints)


@pytest.mark.parametrize("shape,kw,T", CASES)
def test_device_path_matches_oracle(shape, kw, T):
    import torch
    st = synth.make_stream(shape, 7, T, **kw)
    trk = _tracker_for([st])
    dets = torch.from_numpy(st.dets[None]).cuda()
    counts = torch.from_numpy(st.counts[None]).cuda()
    out = trk.run(dets, counts, nviews=True, assoc=True)
    assert trk.check().tolist() == [0]
    out = {k: v.cpu().numpy() for k, v in out.items()}
    oo, oa, _ = util.run_oracle(st)
    worst = util.compare_with_oracle(out, 0, st, oo, oa)
    assert out["count"].sum() > 0
    print(f"{shape}: max joint deviation {worst:.3e} m over {T} frames")


@pytest.mark.parametrize("shape,mult,env", [("shelf", 6.0, ""), ("shelf", 6.0, "4:1:128"), ("shelf", 6.0, "1:8:128"),
                                            ("campus", 6.0, "1:8:128"), ("panoptic", 4.0, "")])
def test_wide_association_threshold_exercises_the_assignment_solver(shape, mult, env, monkeypatch):
    """See tests/test_hostemu_vs_oracle.py: with alpha2d several times the dataset value most frames need the full
    assignment solver, i.e. the values (not only the signs) of the affinities of the contested cameras."""
    import torch
    if env:
        monkeypatch.setenv("PAM_TRACK_SHAPE", env)
    S = 9 if env.startswith("1:") else 2
    streams = [synth.make_stream(shape, 5 + k, 160, miss_prob=0.1, outlier_prob=0.1) for k in range(2)]
    p = synth.tracker_params(shape)
    p["alpha2d"] = p["alpha2d"] * mult
    cams = camera.GetCameraParameters(streams[0].rig)
    trk = tracker.SequenceTracker(cams, p, num_sequences=S, max_detections=streams[0].dets.shape[2],
                                  max_tracks=16 if shape == "panoptic" else 8, arm_joints=streams[0].shape.arm_joints)
    dets = torch.from_numpy(np.stack([streams[k % 2].dets for k in range(S)])).cuda()
    counts = torch.from_numpy(np.stack([streams[k % 2].counts for k in range(S)])).cuda()
    out = trk.run(dets, counts, nviews=True, assoc=True)
    assert trk.check().tolist() == [0] * S
    out = {k: v.cpu().numpy() for k, v in out.items()}
    for s in (0, 1, S - 1):
        oo, oa, _ = util.run_oracle(streams[s % 2], params=p)
        util.compare_with_oracle(out, s, streams[s % 2], oo, oa)


def _valid(out, key):
    """Mask the unused tail of the per-frame output slots (not written by the kernel)."""
    a = out[key]
    if key in ("count", "assoc"):
        return a
    k = out["count"]
    idx = np.arange(a.shape[2])[None, None, :] < k[:, :, None]
    m = idx.reshape(idx.shape + (1,) * (a.ndim - 3))
    return np.where(m, a, 0)


def test_host_path_and_chunked_frames_match():
    """pam_track_sequences_host (H2D + kernel + D2H) equals the device-pointer path, and running a
    sequence in several calls (state carried in HBM between launches) equals one call."""
    import torch
    streams = [synth.make_stream("shelf", 20 + s, 120, miss_prob=0.05, outlier_prob=0.03) for s in range(3)]
    dets = np.stack([s.dets for s in streams])
    counts = np.stack([s.counts for s in streams])
    trk = _tracker_for(streams)
    ref = {k: v.cpu().numpy() for k, v in trk.run(torch.from_numpy(dets).cuda(), torch.from_numpy(counts).cuda(),
                                                  assoc=True).items()}
    trk.check()
    host = trk.run_host(dets, counts, fresh=True, assoc=True)
    for k in ("count", "ids", "joints", "nviews", "assoc"):
        assert np.array_equal(_valid(ref, k), _valid(host, k)), k
    # chunked: 50 + 1 + 69 frames
    trk2 = _tracker_for(streams)
    parts = []
    for a, b in ((0, 50), (50, 51), (51, 120)):
        parts.append(trk2.run_host(np.ascontiguousarray(dets[:, a:b]), np.ascontiguousarray(counts[:, a:b]),
                                   fresh=(a == 0), assoc=True))
    cat = {k: np.concatenate([p[k] for p in parts], 1) for k in ("count", "ids", "joints", "nviews", "assoc")}
    for k in ("count", "ids", "joints", "nviews", "assoc"):
        assert np.array_equal(_valid(ref, k), _valid(cat, k)), k


def test_state_readback_matches_oracle_tracks():
    import torch
    st = synth.make_stream("shelf", 11, 90, miss_prob=0.1, outlier_prob=0.04)
    trk = _tracker_for([st])
    trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda())
    trk.check()
    state = trk.read_state()[0]
    _, _, otrk = util.run_oracle(st)
    assert [t["track_id"] for t in state["tracks"]] == [t.track_id for t in otrk.tracks]
    for got, ref in zip(state["tracks"], otrk.tracks):
        assert (got["hits"], got["age"], got["time_since_update"], got["state"]) == \
               (ref.hits, ref.age, ref.time_since_update, ref.state)
        assert list(got["poses2d"].keys()) == list(ref.poses2d.keys())
        for cid in got["poses2d"]:
            assert got["poses2d"][cid]["time"] == ref.poses2d[cid]["time"]
            assert np.array_equal(got["poses2d"][cid]["pose"], ref.poses2d[cid]["pose"])
        assert [p["time"] for p in got["poses3d"]] == [p["time"] for p in ref.poses3d]
        for a, b in zip(got["poses3d"], ref.poses3d):
            assert np.abs(a["pose3d"] - b["pose3d"]).max() < 1e-6
        assert np.abs(got["velocity_3d"] - ref.velocity_3d).max() < 1e-6


def test_capacity_overflow_is_reported():
    import torch
    st = synth.make_stream("shelf", 3, 30)
    cams = camera.GetCameraParameters(st.rig)
    trk = tracker.SequenceTracker(cams, synth.tracker_params(st.shape), 1, max_detections=4, max_tracks=2,
                                  arm_joints=st.shape.arm_joints)
    trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda())
    with pytest.raises(tracker.PamError) as e:
        trk.check()
    assert "capacity" in str(e.value)


@pytest.mark.parametrize("env", [
    {"PAM_TRACK_SHAPE": "1:8:128"},     # one warp per sequence, 8 sequences per CTA, frames of the CTA started together
    {"PAM_TRACK_SHAPE": "1:12:80"},     # 24 sequences per SM
    {"PAM_TRACK_SHAPE": "1:8:80", "PAM_TRACK_CONVOY": "2"},   # ... every phase together
    {"PAM_TRACK_SHAPE": "1:3:128", "PAM_TRACK_CONVOY": "0"},  # free running, ragged last CTA
    {"PAM_TRACK_SHAPE": "1:8:80:0"},    # one warp per sequence with the latency flavour of the working set
    {"PAM_TRACK_SHAPE": "3:1:80"},      # three warps, one sequence per CTA (round-1 shape)
    {"PAM_TRACK_SHAPE": "3:1:80:1"},    # ... with one detection buffer and the raw pose in HBM scratch
    {"PAM_TRACK_SHAPE": "4:1:128"},     # single-stream default
    {"PAM_TRACK_SHAPE": "8:1:128"},
    # Panoptic-sized working set (12 track slots): capacity class "mid"
    {"PAM_TRACK_SHAPE": "1:6:128", "max_tracks": "12"},
    {"PAM_TRACK_SHAPE": "1:8:80", "max_tracks": "12"},
    {"PAM_TRACK_SHAPE": "3:1:80", "max_tracks": "12"},
    {"PAM_TRACK_SHAPE": "8:1:128", "max_tracks": "12"},
    # largest working set (32 track slots): capacity class "max"
    {"PAM_TRACK_SHAPE": "1:2:128", "max_tracks": "32"},
    {"PAM_TRACK_SHAPE": "4:1:128", "max_tracks": "32"},
])
def test_every_launch_shape_matches_oracle(env, monkeypatch):
    """The launch shape (warps per sequence : sequences per CTA : register budget [: lean working set]) is picked
    from the batch size; small test batches only ever see the single-stream shape, so every other variant is
    forced here (read at pam_create) and checked."""
    import torch
    env = dict(env)
    max_tracks = int(env.pop("max_tracks", "8"))
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    streams = [synth.make_stream("shelf", 40 + s, 150, miss_prob=0.08, outlier_prob=0.04, enter_stagger=10 * s) for s in range(5)]
    trk = _tracker_for(streams, max_tracks=max_tracks)
    dets = torch.from_numpy(np.stack([st.dets for st in streams])).cuda()
    counts = torch.from_numpy(np.stack([st.counts for st in streams])).cuda()
    out = trk.run(dets, counts, nviews=True, assoc=True)
    assert trk.check().tolist() == [0] * 5
    info = trk.launch_info()
    g, q = (int(x) for x in env["PAM_TRACK_SHAPE"].split(":")[:2])
    assert (info["warps_per_sequence"], info["sequences_per_cta"]) == (g, q), info
    out = {k: v.cpu().numpy() for k, v in out.items() if v is not None}
    for s, st in enumerate(streams):
        oo, oa, _ = util.run_oracle(st)
        util.compare_with_oracle(out, s, st, oo, oa)


def test_max_report_only_changes_the_output_stride():
    """pam_config.max_report: fewer output rows per frame (less D2H traffic); the rows kept are the first rows of the
    full output, `count` keeps the true number."""
    import torch
    st = synth.make_stream("shelf", 9, 120, miss_prob=0.05, outlier_prob=0.03)
    cams = camera.GetCameraParameters(st.rig)
    kw = dict(max_detections=st.dets.shape[2], max_tracks=8, arm_joints=st.shape.arm_joints)
    full = tracker.SequenceTracker(cams, synth.tracker_params("shelf"), 1, **kw)
    cut = tracker.SequenceTracker(cams, synth.tracker_params("shelf"), 1, max_report=3, **kw)
    d, c = torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda()
    a = {k: v.cpu().numpy() for k, v in full.run(d, c, assoc=True, vlist=True).items()}
    b = {k: v.cpu().numpy() for k, v in cut.run(d, c, assoc=True, vlist=True).items()}
    assert b["ids"].shape[2] == 3 and b["joints"].shape[2] == 3 and b["nviews"].shape[2] == 3 and b["vlist"].shape[2] == 3
    assert np.array_equal(a["count"], b["count"]) and np.array_equal(a["assoc"], b["assoc"]) and a["count"].max() == 4
    for t in range(st.T):
        r = min(3, a["count"][0, t])
        for k in ("ids", "joints", "nviews", "vlist"):
            assert np.array_equal(a[k][0, t, :r], b[k][0, t, :r]), (k, t)
    h = cut.run_host(st.dets[None], st.counts[None], fresh=True, assoc=True, vlist=True)
    for k in ("count", "assoc"):
        assert np.array_equal(h[k], b[k])
    for t in range(st.T):
        r = min(3, a["count"][0, t])
        for k in ("ids", "joints", "nviews", "vlist"):
            assert np.array_equal(h[k][0, t, :r], b[k][0, t, :r]), (k, t)
