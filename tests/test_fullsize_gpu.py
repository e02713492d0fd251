"""BASELINE.json full-size configurations through the C ABI.

Configs 0 (Campus-shaped, 2000 frames), 1 (Shelf-shaped, 3200 frames) and 2 (Panoptic-shaped, 10 000 frames)
are each compared frame by frame with the oracle over their WHOLE length (one sequence each; the oracles run
in a process pool beside the GPU work).  Configs 1 and 2 are additionally checked through size-independent
properties on a batch: bit-identical results when the frames are fed in uneven chunks, when the run is
repeated, and when a sequence is tracked alone instead of inside a batch; and tracking quality against the
synthetic ground truth."""
import numpy as np
import pytest

from tests import util
from pam_b200 import camera, synth, tracker

pytestmark = pytest.mark.gpu


def _tracker(shape, rig, S, D, max_tracks=12):
    sh = synth.SHAPES[shape]
    return tracker.SequenceTracker(camera.GetCameraParameters(rig), synth.tracker_params(shape), S, max_detections=D,
                                   max_tracks=max_tracks, arm_joints=sh.arm_joints)


def _valid(out, key):
    a = out[key]
    if key == "count":
        return a
    idx = np.arange(a.shape[2])[None, None, :] < out["count"][:, :, None]
    return np.where(idx.reshape(idx.shape + (1,) * (a.ndim - 3)), a, 0)


def _mpjpe(out, gt, s, t0):
    errs, full = [], 0
    P = gt.shape[2]
    for t in range(t0, gt.shape[1]):
        k = out["count"][s, t]
        full += int(k == P)
        for j in range(k):
            d = np.linalg.norm(gt[s, t] - out["joints"][s, t, j][None], axis=2).mean(1)
            errs.append(d.min())
    return float(np.mean(errs)), full / (gt.shape[1] - t0)


def test_config0_campus_2000_frames_full_oracle_parity():
    st = synth.make_stream("campus", 0)          # 3 cameras, 3 people, 14 joints, 2000 frames
    assert st.T == 2000
    trk = _tracker("campus", st.rig, 1, st.dets.shape[2])
    out = trk.run_host(st.dets[None], st.counts[None], fresh=True, assoc=True)
    oo, oa, _ = util.run_oracle(st)
    worst = util.compare_with_oracle(out, 0, st, oo, oa)
    assert worst < 5e-4 and out["count"].sum() > 5000


def _oracle_full(shape):
    """Worker: the oracle over the full length of sequence 0 of a named shape."""
    from tests import util as u
    from pam_b200 import synth as sy
    st = sy.make_stream(shape, 0)
    oo, oa, _ = u.run_oracle(st)
    return shape, oo, oa


@pytest.fixture(scope="module")
def full_length_oracles():
    """Start the full-length oracles of the Shelf and Panoptic configurations in two worker processes right
    away (Panoptic: 10 000 frames x 8 people, about a minute of one core); tests collect them when needed."""
    import multiprocessing as mp
    pool = mp.get_context("spawn").Pool(2)
    pending = {sh: pool.apply_async(_oracle_full, (sh,)) for sh in ("panoptic", "shelf")}
    yield pending
    pool.terminate()


@pytest.mark.parametrize("shape", ["shelf", "panoptic"])
def test_full_length_oracle_parity(shape, full_length_oracles):
    """BASELINE configs 1 and 2 at their stated length: every frame of one sequence against the oracle
    (ids, reported sets, per-joint view counts, associations exact; joints within 0.5 mm / 1e-3)."""
    sh = synth.SHAPES[shape]
    st = synth.make_stream(shape, 0)
    assert st.T == sh.T
    trk = _tracker(shape, st.rig, 1, st.dets.shape[2])
    out = trk.run_host(st.dets[None], st.counts[None], fresh=True, nviews=True, assoc=True)
    assert trk.check(host_path=True).tolist() == [0]
    _, oo, oa = full_length_oracles[shape].get(timeout=900)
    assert len(oo) == sh.T
    worst = util.compare_with_oracle(out, 0, st, oo, oa)
    assert worst < 5e-4 and out["count"].sum() > sh.T * (sh.P - 1)
    print(f"{shape}: {sh.T} frames, max joint deviation {worst:.3e} m")


@pytest.mark.parametrize("shape,S,prefix", [("shelf", 12, 250), ("panoptic", 2, 120)])
def test_full_length_properties(shape, S, prefix):
    sh = synth.SHAPES[shape]
    rig, dets, counts, gt, streams = synth.make_batch(shape, S)
    T = sh.T
    assert dets.shape[1] == T
    trk = _tracker(shape, rig, S, dets.shape[3])
    ref = trk.run_host(dets, counts, fresh=True, nviews=True, assoc=False)
    # oracle parity on a prefix of the first sequence
    oo, oa, _ = util.run_oracle(streams[0], T=prefix)
    pre = {k: (v[:, :prefix] if v is not None else None) for k, v in ref.items()}
    assert util.compare_with_oracle(pre, 0, streams[0], oo, oa) < 5e-4
    # determinism
    again = trk.run_host(dets, counts, fresh=True, nviews=True, assoc=False)
    for k in ("count", "ids", "joints", "nviews"):
        assert np.array_equal(_valid(ref, k), _valid(again, k)), k
    # uneven chunks of frames, tracker state carried in HBM between the calls
    cuts = [0, 1, 2, 9, 137, 1000, 1001, T // 2, T - 3, T]
    parts = [trk.run_host(np.ascontiguousarray(dets[:, a:b]), np.ascontiguousarray(counts[:, a:b]), fresh=(a == 0),
                          nviews=True) for a, b in zip(cuts[:-1], cuts[1:])]
    cat = {k: np.concatenate([p[k] for p in parts], 1) for k in ("count", "ids", "joints", "nviews")}
    for k in ("count", "ids", "joints", "nviews"):
        assert np.array_equal(_valid(ref, k), _valid(cat, k)), k
    # a sequence tracked alone gives the same result as inside the batch
    solo = _tracker(shape, rig, 1, dets.shape[3])
    one = solo.run_host(dets[S - 1:S], counts[S - 1:S], fresh=True, nviews=True)
    for k in ("count", "ids", "joints", "nviews"):
        assert np.array_equal(_valid(ref, k)[S - 1:S], _valid(one, k)), k
    # tracking quality against the synthetic ground truth
    for s in range(S):
        err, full = _mpjpe(ref, gt, s, 10)
        assert err < 0.02 and full > 0.97, (s, err, full)
        ids = ref["ids"][s][np.arange(ref["ids"].shape[2])[None, :] < ref["count"][s][:, None]]
        assert len(set(ids.tolist())) <= sh.P + 3
