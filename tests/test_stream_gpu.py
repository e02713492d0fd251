"""Stream mode (pam_stream_*): the resident per-frame kernel behind IterativeTracker.tracking().  One frame per
call, tracker state on chip; results must equal the oracle's frame by frame, also across a state read-back
(which parks the kernel), an idle period (the kernel leaves after about a second and is restarted) and a restart."""
import time

import numpy as np
import pytest

from tests import util
from pam_b200 import camera, synth, tracker

pytestmark = pytest.mark.gpu


def _check_frame(fs, t, oo, oa):
    ids, joints, views = oo[t]
    k = int(fs.count[0])
    assert k == len(ids), (t, k, len(ids))
    assert np.array_equal(fs.ids[:k], ids), t
    if k:
        assert np.array_equal(fs.nviews[:k], views), t
        err = np.abs(fs.joints[:k].astype(np.float64) - joints)
        assert np.all(err <= np.maximum(5e-4, 1e-3 * np.abs(joints))), (t, err.max())
    for c, a in enumerate(oa[t]):
        assert np.array_equal(fs.assoc[c, :len(a)], a), (t, c)


@pytest.mark.timeout(120)
def test_stream_mode_matches_oracle_frame_by_frame():
    st = synth.make_stream("shelf", 31, 220, miss_prob=0.08, outlier_prob=0.04, enter_stagger=15, absences=[(2, 90, 120)])
    cams = camera.GetCameraParameters(st.rig)
    trk = tracker.SequenceTracker(cams, synth.tracker_params("shelf"), 1, max_detections=st.dets.shape[2], max_tracks=8,
                                  arm_joints=st.shape.arm_joints)
    oo, oa, otrk = util.run_oracle(st)
    fs = trk.open_stream(fresh=True)
    for t in range(st.T):
        fs.set_padded(st.dets[t], st.counts[t])
        fs.step(t)
        _check_frame(fs, t, oo, oa)
        assert fs.timing[3] > 0 and fs.timing[0] + fs.timing[1] + fs.timing[2] <= fs.timing[3]
        if t == 100:
            # state read-back parks the resident kernel (the state lives on chip while it runs) ...
            state = trk.read_state(host_path=True)[0]
            assert [x["track_id"] for x in state["tracks"]]
        if t == 150:
            time.sleep(1.6)      # ... and so does an idle second; the next step restarts it
    state = trk.read_state(host_path=True)[0]
    assert [x["track_id"] for x in state["tracks"]] == [x.track_id for x in otrk.tracks]
    for got, ref in zip(state["tracks"], otrk.tracks):
        assert (got["hits"], got["age"], got["time_since_update"], got["state"]) == (ref.hits, ref.age, ref.time_since_update, ref.state)
        assert list(got["poses2d"].keys()) == list(ref.poses2d.keys())
        for cid in got["poses2d"]:
            assert np.array_equal(got["poses2d"][cid]["pose"], ref.poses2d[cid]["pose"])
    # restart: a fresh stream tracks the same frames to the same result
    fs = trk.open_stream(fresh=True)
    for t in range(40):
        fs.set_padded(st.dets[t], st.counts[t])
        fs.step(t)
        _check_frame(fs, t, oo, oa)
    fs.close()
    trk.close()


@pytest.mark.timeout(120)
def test_stream_mode_and_host_path_interoperate():
    """Frames 0-59 through the resident kernel, 60-119 through pam_track_sequences_host (which parks it), 120-179
    through the resident kernel again: one continuous sequence."""
    st = synth.make_stream("panoptic", 4, 180, miss_prob=0.05, outlier_prob=0.03)
    cams = camera.GetCameraParameters(st.rig)
    trk = tracker.SequenceTracker(cams, synth.tracker_params("panoptic"), 1, max_detections=st.dets.shape[2], max_tracks=12,
                                  arm_joints=st.shape.arm_joints)
    oo, oa, _ = util.run_oracle(st)
    fs = trk.open_stream(fresh=True)
    for t in range(60):
        fs.set_padded(st.dets[t], st.counts[t])
        fs.step(t)
        _check_frame(fs, t, oo, oa)
    out = trk.run_host(st.dets[None, 60:120], st.counts[None, 60:120], frame0=60, assoc=True)
    util.compare_with_oracle(out, 0, st, oo[60:120], oa[60:120])
    fs = trk.open_stream(fresh=False)
    for t in range(120, 180):
        fs.set_padded(st.dets[t], st.counts[t])
        fs.step(t)
        _check_frame(fs, t, oo, oa)
    trk.close()


@pytest.mark.timeout(120)
def test_dropin_tracking_returns_device_timers_and_guards_float32():
    D = util.load_dropin()
    from types import SimpleNamespace
    st = synth.make_stream("shelf17", 3, 30)
    cams = camera.GetCameraParameters(st.rig)
    trk = D.IterativeTracker.IterativeTracker(SimpleNamespace(**synth.tracker_params("shelf17")))
    for t in range(st.T):
        a, u, i = trk.tracking(t, cams, [None] * 5, st.frame_boxes(t), st.frame_detections(t), "SVD")
        assert a > 0 and u > 0 and i >= 0 and a + u + i < 1e-3      # seconds on the device, a few microseconds each
    dets = st.frame_detections(5)
    dets[2][0, 3, 0] += 1e-9                      # not representable in float32
    with pytest.raises(ValueError):
        trk.tracking(st.T, cams, [None] * 5, st.frame_boxes(5), dets, "SVD")
    D.IterativeTracker.IterativeTracker.STRICT_FLOAT32 = False
    try:
        trk.tracking(st.T, cams, [None] * 5, st.frame_boxes(5), dets, "SVD")
    finally:
        D.IterativeTracker.IterativeTracker.STRICT_FLOAT32 = True
