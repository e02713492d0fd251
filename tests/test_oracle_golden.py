"""The oracle (oracle/generic.py) against the golden vectors generated from the UNMODIFIED
reference (tests/golden/make_golden.py), and -- where /root/reference exists -- bit-for-bit against
the reference itself.  CPU only."""
import numpy as np
import pytest

from tests import util
from oracle import generic, ref_loader
from pam_b200 import synth


def _cams(rig):
    return generic.build_cameras(rig["P"], rig["K"], rig["RT"])


@pytest.mark.parametrize("n", [0, 1, 2])
def test_oracle_reproduces_golden_streams(n):
    st, g = util.golden_streams()[n]
    out, assoc, _ = generic.run_stream(st, synth.tracker_params(st.shape), (9, 10), 10, trace=True)
    for t, (ids, joints, views) in enumerate(out):
        k = g["count"][t]
        assert len(ids) == k and np.array_equal(ids, g["ids"][t, :k]), t
        assert np.array_equal(views, g["views"][t, :k]), t
        assert np.allclose(joints, g["joints"][t, :k], rtol=0, atol=1e-9), t
        for c, a in enumerate(assoc[t]):
            assert np.array_equal(a, g["assoc"][t, c, :len(a)]), (t, c)


def test_oracle_functions_match_golden():
    z = util.golden_functions()
    cams = _cams(dict(P=z["P"], K=z["K"], RT=z["RT"]))
    assert np.array_equal(np.stack([c.F for c in cams]), z["cam_F"])
    assert np.array_equal(np.stack([c.RK_INV for c in cams]), z["cam_RK_INV"])
    assert np.array_equal(np.stack([c.position for c in cams]), z["cam_position"])
    tol = dict(rtol=0, atol=1e-9)
    assert np.allclose(np.stack([c.project_tracks(z["proj_in"]) for c in cams]), z["proj_out"], **tol)
    sub = [cams[i] for i in z["eap_order"]]
    m, D = generic.epipolar_affinity_parallel(sub, np.arange(len(sub)), z["eap_pose"], 17)
    assert np.allclose(m, z["eap_mean"], **tol) and np.allclose(D, z["eap_D"], **tol)
    m, D = generic.epipolar_affinity(cams, z["ea_cam"], z["ea_pose"], 17)
    assert m.dtype == np.float32 and np.allclose(m, z["ea_mean"], rtol=0, atol=1e-4) and np.allclose(D, z["ea_D"], rtol=0, atol=1e-4)
    assert np.allclose(generic.epipolar_distance(cams[1], z["pa"][1], cams[3], z["pb"][3]), z["ed_out"], **tol)
    for b in range(len(z["gm_A"])):
        _, keep, _ = generic.greedy_view_filter(sub, pose_mat=z["gm_pose"][b].reshape(-1, 1, 3), affinity_mat=z["gm_A"][b],
                                                next_pose=z["gm_next"][b])
        assert np.array_equal(keep[::2], z["gm_keep_update"][b])
        _, keep, _ = generic.greedy_view_filter(sub, affinity_mat=z["gm_A"][b].astype(np.float32), mode="init")
        assert np.array_equal(keep[::2], z["gm_keep_init"][b])
    keep = z["svd_keep"]
    jv = [[] for _ in sub]
    for j in range(17):
        jv[int(keep[j, ::2].sum()) - 1].append(j)
    got = generic.dlt_joint_filtered(sub, list(z["svd_Ts"]), z["eap_pose"], 5, keep, jv, z["svd_next"])
    assert np.allclose(got, z["svd_jf"], rtol=0, atol=1e-7)
    assert np.allclose(generic.dlt_all_views(sub, list(z["svd_Ts"]), z["eap_pose"], 5), z["svd_parallel"], rtol=0, atol=1e-7)
    dirs = generic.pixel_rays(cams[2].RK_INV, cams[2].position, z["ray_uv"])
    assert np.allclose(dirs, z["ray_dirs"], **tol)
    assert np.allclose(generic.ray_point_distance(cams[2].position, dirs, z["ray_X"]), z["ray_dist"], **tol)
    hyp = generic.Hypothesis(cams[0], z["pa"][0], 60)
    hyp.merge(cams[2], z["pa"][2])
    c1, v1 = hyp.calculate_cost(cams[3], z["pa"][3])
    c2, v2 = hyp.calculate_cost(cams[3], z["pb"][3])
    assert np.allclose([c1, c2], z["hyp_cost"], **tol) and [v1, v2] == list(z["hyp_veto"])
    hyp.merge(cams[3], z["pa"][3])
    hyp.merge(cams[4], z["pa"][4])
    _, _, p3d, _, ok = hyp.get_3dpose_jf(30, 5)
    assert ok == bool(z["hyp_ok"]) and np.allclose(p3d, z["hyp_pose3d"], rtol=0, atol=1e-7)
    assert np.allclose([generic.mean_confidence(z["pa"][0]), generic.mean_confidence(z["pb"][1])], z["believe"], **tol)


def test_epilines_restatement_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for _ in range(5):
        F = (rng.normal(size=(3, 3)) * 10 ** rng.uniform(0, 5)).astype(np.float32)
        pts = rng.uniform(0, 1500, size=(17, 2))
        for which in (1, 2):
            ref = np.squeeze(cv2.computeCorrespondEpilines(pts, which, F))
            assert np.array_equal(ref, generic.epilines(pts, which, F))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("shape,kw", [("shelf17", dict(enter_stagger=20, miss_prob=0.1, outlier_prob=0.05)),
                                      ("campus17", dict(miss_prob=0.05, outlier_prob=0.03, absences=[(0, 30, 50)]))])
def test_oracle_bit_identical_to_unmodified_reference(shape, kw):
    import warnings
    warnings.filterwarnings("ignore")
    st = synth.make_stream(shape, 3, 120, **kw)
    V = st.shape.V
    rcams = ref_loader.make_cameras(st.rig["P"], st.rig["K"], st.rig["RT"])
    ocams = _cams(st.rig)
    for a, b in zip(rcams, ocams):
        for k in ("P", "K", "RT", "F", "RK_INV", "position"):
            assert np.array_equal(getattr(a, k), getattr(b, k)) and getattr(a, k).dtype == getattr(b, k).dtype
    prm = synth.tracker_params(shape)
    rtrk, otrk = ref_loader.make_tracker(prm), generic.Tracker(prm, (9, 10), 10)
    for t in range(st.T):
        rtrk.tracking(t, rcams, [None] * V, st.frame_boxes(t), st.frame_detections(t), "SVD")
        otrk.tracking(t, ocams, [None] * V, st.frame_boxes(t), st.frame_detections(t), "SVD")
        assert len(rtrk.tracks) == len(otrk.tracks)
        for x, y in zip(rtrk.tracks, otrk.tracks):
            assert (x.track_id, x.state, x.hits, x.time_since_update) == (y.track_id, y.state, y.hits, y.time_since_update)
            assert list(x.poses2d.keys()) == list(y.poses2d.keys())
            assert np.array_equal(x.poses3d[-1]["pose3d"], y.poses3d[-1]["pose3d"])
            assert x.poses3d[-1]["joints_views"] == y.poses3d[-1]["joints_views"]
            assert np.array_equal(x.velocity_3d, y.velocity_3d) and x.velocity_3d.dtype == y.velocity_3d.dtype
