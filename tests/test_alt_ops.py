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
