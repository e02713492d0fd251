"""Stateless batched ops against the oracle at the joint counts / camera counts the unmodified
reference cannot run (J = 14, 19; the 31-camera dense crowd), plus edge cases."""
import numpy as np
import pytest

from tests import util
from oracle import generic
from pam_b200 import camera, ops, synth

pytestmark = pytest.mark.gpu


def _rig(shape):
    rig = synth.make_rig(shape)
    return camera.GetCameraParameters(rig), generic.build_cameras(rig["P"], rig["K"], rig["RT"])


def test_dense_allpairs_affinity_and_triangulation_match_oracle():
    """config 3: 31 cameras, 19 joints; a 6-person slice keeps the O(M^2) python oracle in seconds."""
    st = synth.make_stream("dense", 0, 1, P=6, miss_prob=0.0, outlier_prob=0.02)
    cams, ocams = _rig("dense")
    V, J = st.shape.V, st.shape.J
    poses = np.concatenate([st.dets[0, c, :st.counts[0, c]].astype(np.float64) for c in range(V)])
    cam_idx = np.concatenate([np.full(st.counts[0, c], c) for c in range(V)])
    o = ops.GeometryOps(cams, J, synth.tracker_params("dense"))
    aff, D = o.epipolar_allpairs(cam_idx, poses)
    raff, rD = generic.epipolar_affinity(ocams, cam_idx, poses, J)
    assert aff.dtype == np.float32 and np.allclose(aff, raff, rtol=1e-6, atol=1e-4)
    assert np.allclose(D, rD, rtol=1e-6, atol=1e-4)
    aff2, none = o.epipolar_allpairs(cam_idx, poses, want_dist=False)
    assert none is None and np.array_equal(aff, aff2)
    # triangulate every person from all 31 views (ground-truth grouping), random ages
    rng = np.random.default_rng(1)
    P = 6
    pm = np.zeros((P, V, J, 3))
    for c in range(V):
        for d in range(st.counts[0, c]):
            pm[st.person_of_det[0, c, d], c] = st.dets[0, c, d]
    Ts = rng.choice([0, 0, 0, 1, 2], size=(P, V))
    got = o.triangulate(np.tile(np.arange(V), (P, 1)), pm, np.exp(-5.0 * Ts))
    for p in range(P):
        ref = generic.dlt_all_views(ocams, list(Ts[p]), pm[p], 5)
        assert np.abs(got[p] - ref).max() < 5e-4
        assert np.abs(got[p] - st.gt[0, p]).max() < 0.25      # and it is the right person (outlier joints included)


def test_dense_full_size_sampled_pairs_and_all_triangulations():
    """config 3 at its stated size: 31 cameras x 64 people x 19 joints, M = 1984 detections.  The O(M^2) python
    oracle cannot fill the whole (M, M, J) tensor in test time, so 2000 random (i, j) pairs (plus same-camera and
    diagonal entries) of the device result are compared with generic.epipolar_distance stored as float32 like the
    reference (utils/matching.py:96-112), and ALL 64 x 19 triangulations from 31 views with the oracle's DLT."""
    st = synth.make_stream("dense", 0, 1, miss_prob=0.0, outlier_prob=0.01)
    cams, ocams = _rig("dense")
    V, J, P = st.shape.V, st.shape.J, st.shape.P
    poses = np.concatenate([st.dets[0, c, :st.counts[0, c]].astype(np.float64) for c in range(V)])
    cam_idx = np.concatenate([np.full(st.counts[0, c], c) for c in range(V)])
    M = len(poses)
    assert M == V * P == 1984
    o = ops.GeometryOps(cams, J, synth.tracker_params("dense"))
    aff, D = o.epipolar_allpairs(cam_idx, poses)
    assert aff.shape == (M, M) and D.shape == (M, M, J) and aff.dtype == np.float32
    aff2, _ = o.epipolar_allpairs(cam_idx, poses, want_dist=False)
    assert np.array_equal(aff, aff2)
    rng = np.random.default_rng(7)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, M, size=(2000, 2))]
    pairs += [(k, k) for k in (0, 1, M - 1)] + [(0, 1), (P, P + 5), (M - 2, M - 1)]      # diagonal + same-camera pairs
    worst = 0.0
    for a, b in pairs:
        if a == b:
            assert aff[a, b] == 0.0 and not D[a, b].any()
            continue
        if cam_idx[a] == cam_idx[b]:
            assert aff[a, b] == np.float32(25.0)          # the reference's initial value, never overwritten
            continue
        a, b = min(a, b), max(a, b)                         # the reference evaluates the pair as (i < j)
        d = generic.epipolar_distance(ocams[cam_idx[a]], poses[a], ocams[cam_idx[b]], poses[b])
        sym = [(x[0] + x[1]) / 2 for x in d]
        ref = np.asarray(sym, dtype=np.float32)             # stored into a float32 array by the reference
        assert np.allclose(D[a, b], ref, rtol=1e-6, atol=1e-4), (a, b)
        assert np.array_equal(D[a, b], D[b, a])
        assert np.isclose(aff[a, b], np.float32(np.mean(sym)), rtol=1e-6, atol=1e-4)
        assert aff[a, b] == aff[b, a]
        worst = max(worst, float(np.abs(D[a, b] - ref).max()))
    # all 64 x 19 triangulations from 31 views, ground-truth grouping, random ages
    pm = np.zeros((P, V, J, 3))
    for c in range(V):
        for d in range(st.counts[0, c]):
            pm[st.person_of_det[0, c, d], c] = st.dets[0, c, d]
    Ts = rng.choice([0, 0, 0, 1, 2], size=(P, V))
    got = o.triangulate(np.tile(np.arange(V), (P, 1)), pm, np.exp(-5.0 * Ts))
    dmax = 0.0
    for p in range(P):
        ref = generic.dlt_all_views(ocams, list(Ts[p]), pm[p], 5)
        dmax = max(dmax, float(np.abs(got[p] - ref).max()))
    assert dmax < 5e-4
    print(f"dense M={M}: {len(pairs)} sampled pairs, max |dD| {worst:.2e} px; {P * J} triangulations, max |dX| {dmax:.2e} m")


@pytest.mark.parametrize("shape", ["campus", "panoptic"])
def test_per_track_ops_match_oracle(shape):
    st = synth.make_stream(shape, 5, 3, miss_prob=0.0, outlier_prob=0.1)
    cams, ocams = _rig(shape)
    V, J, P = st.shape.V, st.shape.J, st.shape.P
    prm = synth.tracker_params(shape)
    o = ops.GeometryOps(cams, J, prm)
    # association affinity of "tracks" = ground truth of frame 0 against detections of frame 1
    tracks = st.gt[0]
    dt = np.array([1 + (i % 3) for i in range(P)])
    dets = st.frame_detections(1)
    got = o.assoc_affinity(tracks, dt, dets)
    for c in range(V):
        reproj = ocams[c].project_tracks(tracks)
        n, m = P, len(dets[c])
        c2d = np.linalg.norm(np.repeat(reproj, m, 0) - np.tile(dets[c][:, :, :2], (n, 1, 1)), axis=2).reshape(n, m, -1)
        c2d = 1 - np.transpose(c2d.T / (prm["alpha2d"] * dt))
        with np.errstate(all="ignore"):
            aff = np.sum(c2d, where=c2d > 0, axis=2) / np.sum(c2d > 0, axis=2)
        aff[~(np.sum(c2d > 0, axis=2) > 10)] = 0
        aff = np.transpose(aff.T / np.exp(prm["lambda_a"] * dt))
        aff[np.isnan(aff)] = 0
        assert np.allclose(got[c], aff, rtol=1e-9, atol=1e-12)
        # the assignment agrees with scipy on the accepted (positive) pairs
        from scipy.optimize import linear_sum_assignment
        r, cc = linear_sum_assignment(-aff)
        r2, c2 = o.assign(-got[c])
        assert {(a, b) for a, b in zip(r, cc) if aff[a, b] > 0} == {(a, b) for a, b in zip(r2, c2) if aff[a, b] > 0}
    # per-track epipolar distances in a scrambled view order
    order = list(np.random.default_rng(0).permutation(V))
    person = 1
    pm = np.array([st.dets[1, c, list(st.person_of_det[1, c]).index(person)].astype(np.float64) for c in order])
    sub, osub = [cams[c] for c in order], [ocams[c] for c in order]
    o2 = ops.GeometryOps(sub, J, prm)
    mean, D = o2.epipolar_pairs(np.arange(V), pm)
    rmean, rD = generic.epipolar_affinity_parallel(osub, np.arange(V), pm, J)
    assert np.allclose(D, rD, rtol=1e-9, atol=1e-7) and np.allclose(mean, rmean, rtol=1e-9, atol=1e-7)
    # view filter + DLT for every joint, against the oracle's per-joint loop
    A = 1 - rD / prm["joint_threshold"]
    nxt = st.gt[1, person] + 0.01
    uv = np.flip(pm[:, :, :2], axis=2)
    keep = o2.view_filter(np.arange(V), np.ascontiguousarray(np.transpose(A, (2, 0, 1))), "update",
                          np.ascontiguousarray(np.transpose(uv, (1, 0, 2))), nxt)
    okeep = np.ones((J, 2 * V), dtype=int)
    jv = [[] for _ in range(V)]
    for j in range(J):
        alive, okeep[j], _ = generic.greedy_view_filter(osub, pose_mat=pm[:, j].reshape(-1, 1, 3), affinity_mat=A[:, :, j],
                                                        next_pose=nxt[j])
        jv[len(alive) - 1].append(j)
    assert np.array_equal(keep, okeep[:, ::2])
    Ts = [0, 1, 0, 2, 0][:V]
    got3d = o2.triangulate(np.arange(V), pm, np.exp(-prm["lambda_t"] * np.array(Ts)), keep=keep[None], next_pose=nxt[None])
    ref3d = generic.dlt_joint_filtered(osub, Ts, pm, prm["lambda_t"], okeep, jv, nxt)
    assert np.abs(got3d - ref3d).max() < 5e-4


def test_assign_matches_scipy_on_random_and_degenerate_matrices():
    from scipy.optimize import linear_sum_assignment
    cams, _ = _rig("campus")
    o = ops.GeometryOps(cams, 14)
    rng = np.random.default_rng(3)
    for nr, nc in [(1, 1), (3, 5), (5, 3), (8, 8), (17, 9), (64, 64), (40, 64)]:
        C = rng.uniform(0, 1, (6, nr, nc))
        C[1] = np.round(C[1] * 4) / 4          # many ties
        C[2][C[2] < 0.7] = 0.0                 # mostly zeros, like an affinity matrix
        res = o.assign(C)
        for b in range(6):
            r, c = linear_sum_assignment(C[b])
            r2, c2 = res[b]
            assert len(r2) == min(nr, nc) and len(set(c2)) == len(c2)
            assert abs(C[b][r, c].sum() - C[b][r2, c2].sum()) < 1e-12
        rm, cm = o.assign(C[0], maximize=True)
        r, c = linear_sum_assignment(C[0], maximize=True)
        assert abs(C[0][r, c].sum() - C[0][rm, cm].sum()) < 1e-12
    assert o.assign(np.zeros((0, 4)))[0].size == 0


def test_empty_and_ragged_inputs():
    cams, ocams = _rig("shelf")
    o = ops.GeometryOps(cams, 14, synth.tracker_params("shelf"))
    st = synth.make_stream("shelf", 2, 2)
    dets = st.frame_detections(1)
    dets[2] = dets[2][:0]                       # a camera without detections
    dets[4] = dets[4][:1]
    got = o.assoc_affinity(st.gt[0], np.ones(4, int), dets)
    assert got[2].shape == (4, 0) and got[4].shape == (4, 1)
    d, _ = o.ray_distance(0, np.zeros((0, 2)), np.zeros((0, 3)))
    assert d.shape == (0,)
