"""alternatives the reference keeps in its API without a live caller -- the One-Euro
smoother (src/tracking/OneEuroFilter.py) and the pair-wise triangulation top_down_pose_kernel
(src/utils/construction.py:9-31).  CPU: the oracle restatements against the UNMODIFIED reference modules, bit for bit.
GPU: libpam's kernels (pam_one_euro, pam_top_down, through the drop-in modules) against the oracle."""
import os

import numpy as np
import pytest

from tests import util
from oracle import filters
from pam_b200 import camera, synth


def _signal(n, seed):
    rng = np.random.default_rng(seed)
    t, xs, ts = 0.0, [], []
    for k in range(n):
        t += 0.04 if k % 7 else float(rng.uniform(0.02, 0.06))
        xs.append(float(np.sin(0.11 * k) + rng.normal(0, 0.05)))
        ts.append(None if k % 11 == 5 else (0.0 if k == 0 else t))        # missing and zero time stamps like :64
    return xs, ts


def _person_views(shape="shelf17", person=0, frame=1):
    st = synth.make_stream(shape, 1, 3, noise_px=1.5)
    pm = []
    for c in range(st.shape.V):
        d = list(st.person_of_det[frame, c]).index(person)
        pm.append(st.dets[frame, c, d, :, :2][:, ::-1].astype(np.float64))     # (x, y) for cv2.triangulatePoints
    return st, pm


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference tree (build container only)")
def test_oracle_restatements_equal_the_unmodified_reference():
    from oracle import ref_loader
    ns = ref_loader.load()
    xs, ts = _signal(400, 0)
    ref = ns.OneEuroFilter.OneEuroFilter(freq=25, mincutoff=0.8, beta=0.4, dcutoff=0.4)
    st = filters.OneEuroState(25, 0.8, 0.4, 0.4)
    for x, t in zip(xs, ts):
        assert ref(x, t) == filters.one_euro_step(st, x, t)
    strm, pm = _person_views()
    cams = ref_loader.make_cameras(strm.rig["P"], strm.rig["K"], strm.rig["RT"])
    w = [0.9, 0.8, 0.7, 0.95, 0.85]
    a, wa = ns.construction.top_down_pose_kernel(cams, pm, w)
    b, wb = filters.top_down_pose_kernel(cams, pm, w)
    assert np.array_equal(a, b) and wa == wb


@pytest.mark.gpu
def test_one_euro_bank_is_bit_identical_to_the_python_class():
    D = util.load_dropin()
    n = 51                                                     # 17 joints x 3 coordinates, common time stamps
    sig = [_signal(300, s) for s in range(n)]
    ts = sig[0][1]
    bank = D.OneEuroFilter.OneEuroBank(n, freq=25, mincutoff=0.8, beta=0.4, dcutoff=0.4)
    states = [filters.OneEuroState(25, 0.8, 0.4, 0.4) for _ in range(n)]
    for k, t in enumerate(ts):
        x = np.array([sig[c][0][k] for c in range(n)])
        got = bank(x, t)
        ref = np.array([filters.one_euro_step(states[c], float(x[c]), t) for c in range(n)])
        assert np.array_equal(got, ref), k
    one = D.OneEuroFilter.OneEuroFilter(freq=120, mincutoff=1.0, beta=1.0, dcutoff=1.0)      # the scalar call surface
    st = filters.OneEuroState(120, 1.0, 1.0, 1.0)
    for x, t in zip(*_signal(60, 99)):
        assert one(x, t) == filters.one_euro_step(st, x, t)
    assert one(None) is None
    with pytest.raises(ValueError):
        D.OneEuroFilter.OneEuroFilter(freq=0)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["shelf17", "panoptic"])
def test_top_down_pose_kernel_matches_oracle(shape):
    from oracle import generic
    D = util.load_dropin()
    strm, pm = _person_views(shape)
    cams = camera.GetCameraParameters(strm.rig)
    ocams = generic.build_cameras(strm.rig["P"], strm.rig["K"], strm.rig["RT"])
    w = list(np.linspace(0.7, 0.95, len(pm)))
    ref, wref = filters.top_down_pose_kernel(ocams, pm, w)
    got, wgot = D.construction.top_down_pose_kernel(cams, pm, w)
    assert wgot == wref                                        # the same pair won
    assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-6
    # all pairs: the summed reprojection errors the argmin is taken over
    from pam_b200 import ops
    o = ops.get_ops(cams, strm.shape.J)
    _, pair, err = o.top_down(np.arange(len(pm)), np.asarray(pm), want_errors=True)
    import cv2
    k = 0
    for i in range(len(pm)):
        for j in range(i + 1, len(pm)):
            homo = cv2.triangulatePoints(ocams[i].P, ocams[j].P, pm[i].T, pm[j].T)
            e = 0.0
            for cam, pk in zip(ocams, pm):
                ph = cam.P @ homo
                e += np.linalg.norm((ph[:2] / (ph[2] + 10e-6)).T - pk)
            assert abs(err[k] - e) < 1e-3 * max(1.0, e), (i, j, err[k], e)      # the "+ 10e-6" depends on cv2's sign / scale of the homogeneous vector
            k += 1
    assert tuple(pair) == min(((i, j) for i in range(len(pm)) for j in range(i + 1, len(pm))),
                              key=lambda ij: err[[(a, b) for a in range(len(pm)) for b in range(a + 1, len(pm))].index(ij)])


class _RecordingSolver:
    """Stands in for the caller's solver object: keeps the matrix it is handed and clusters by a plain threshold."""

    def solve(self, affinity_matrix, rtn_matrix=False):
        self.seen = np.array(affinity_matrix)
        n = len(affinity_matrix)
        labels = list(range(n))
        for i in range(n):
            for j in range(i + 1, n):
                if affinity_matrix[i, j] > 0.5:
                    labels[j] = labels[i]
        return [[k for k in range(n) if labels[k] == lab] for lab in sorted(set(labels))]


@pytest.mark.gpu
def test_bip_matching_front_end_matches_oracle():
    """BIP_matching (src/utils/matching.py:234-241): the matrix handed to the caller's solver and the camera index per
    detection; the solver itself is the caller's object."""
    from oracle import generic
    D = util.load_dropin()
    st = synth.make_stream("shelf17", 2, 2, miss_prob=0.1)
    cams = camera.GetCameraParameters(st.rig)
    ocams = generic.build_cameras(st.rig["P"], st.rig["K"], st.rig["RT"])
    V = st.shape.V
    poses = np.concatenate([st.dets[1, c, :st.counts[1, c]].astype(np.float64) for c in range(V)])
    dim_group = np.concatenate([[0], np.cumsum(st.counts[1])]).astype(int)
    model = _RecordingSolver()
    matched, cam_of = D.matching.BIP_matching(model, cams, dim_group, poses, 17, 40)
    ref_cam = np.concatenate([np.full(st.counts[1, c], c) for c in range(V)]).astype(np.int32)
    assert cam_of.dtype == np.int32 and np.array_equal(cam_of, ref_cam)
    ref_aff, _ = generic.epipolar_affinity(ocams, ref_cam, poses, 17)
    ref_mat = (1 - ref_aff / 40).astype(np.double)
    assert model.seen.dtype == np.float64 and np.allclose(model.seen, ref_mat, rtol=1e-5, atol=1e-5)
    assert matched == _RecordingSolver().solve(ref_mat)
    # detections of one person end up in one cluster (the same person's poses are epipolar-consistent)
    person = np.concatenate([st.person_of_det[1, c, :st.counts[1, c]] for c in range(V)])
    for cl in matched:
        assert len(set(person[cl].tolist())) == 1


@pytest.mark.gpu
def test_kalman_bank_matches_cv2_kalman_filter():
    """KalmanFilter (src/tracking/KalmanFilter.py:4-65): the reference's class is cv2.KalmanFilter(9, 3) with fixed
    matrices; the device bank must follow it through corrections, missing measurements and pure predictions to float32
    rounding (OpenCV inverts the innovation covariance by SVD, the kernel directly)."""
    import cv2
    D = util.load_dropin()

    def reference_filter(pt3d, Hz=25):            # the constructor of KalmanFilter.py:5-50, same cv2 object
        dt = 1.0 / Hz
        v, a = dt, 0.5 * (dt ** 2)
        k = cv2.KalmanFilter(9, 3, 0)
        A = np.eye(9, dtype=np.float32)
        for i in range(6):
            A[i, i + 3] = v
        for i in range(3):
            A[i, i + 6] = a
        k.transitionMatrix = A
        k.measurementMatrix = A[:3].copy()
        k.processNoiseCov = np.eye(9, dtype=np.float32) * 0.007
        k.measurementNoiseCov = np.eye(3, dtype=np.float32) * 0.1
        k.statePre = np.array([[np.float32(pt3d[0])], [np.float32(pt3d[1])], [np.float32(pt3d[2])]] + [[np.float32(0.)]] * 6)
        return k

    def reference_predict(k, pt3d=None):          # KalmanFilter.py:52-65
        if pt3d is not None:
            k.correct(np.array([[np.float32(pt3d[0])], [np.float32(pt3d[1])], [np.float32(pt3d[2])]]))
        return k.predict()[:3].flatten()

    rng = np.random.default_rng(3)
    J, T = 17, 80
    t = np.arange(T)[:, None, None]
    path = rng.uniform(-2, 2, (1, J, 3)) + 0.03 * t * rng.uniform(-1, 1, (1, J, 3)) + 0.002 * rng.normal(size=(T, J, 3))
    bank = D.KalmanFilter.KalmanBank(path[0])
    refs = [reference_filter(path[0, j]) for j in range(J)]
    worst = 0.0
    for k in range(T):
        meas = None if k % 9 == 4 else path[k]                       # a frame without measurements: predict only
        got = bank.predict(meas)
        ref = np.array([reference_predict(refs[j], None if meas is None else meas[j]) for j in range(J)], dtype=np.float64)
        err = np.abs(got - ref).max()
        worst = max(worst, err)
        assert err < 2e-4 * max(1.0, np.abs(ref).max()), (k, err)
    one = D.KalmanFilter.KalmanFilter(path[0, 0])                    # the per-joint call surface
    r1 = reference_filter(path[0, 0])
    for k in range(10):
        a, b = one.predict(path[k, 0]), reference_predict(r1, path[k, 0])
        assert a.dtype == np.float32 and np.abs(a - b).max() < 2e-4
    print(f"kalman bank: max deviation from cv2.KalmanFilter {worst:.2e} over {T} steps x {J} joints")
