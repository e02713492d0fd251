"""CUDA path against the golden vectors produced by the unmodified reference (J = 17)."""
import numpy as np
import pytest

from tests import util
from pam_b200 import camera, synth, tracker

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2])
def test_tracker_kernel_reproduces_reference_streams(n):
    import torch
    st, g = util.golden_streams()[n]
    cams = camera.GetCameraParameters(st.rig)
    trk = tracker.SequenceTracker(cams, synth.tracker_params(st.shape), 1, max_detections=st.dets.shape[2],
                                  max_tracks=12, arm_joints=(9, 10))
    out = trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda(), assoc=True)
    assert trk.check().tolist() == [0]
    out = {k: v.cpu().numpy() for k, v in out.items()}
    frames, assoc = util.golden_as_oracle_lists(g)
    worst = util.compare_with_oracle(out, 0, st, frames, assoc)
    assert worst < 5e-4


def test_dropin_tracker_per_frame_api_reproduces_reference_stream():
    """IterativeTracker.tracking() drop-in, one frame per call, with the reference's read surface."""
    D = util.load_dropin()
    from oracle.ref_loader import EasyDict
    st, g = util.golden_streams()[0]
    cams = camera.GetCameraParameters(st.rig)
    trk = D.IterativeTracker.IterativeTracker(EasyDict(synth.tracker_params(st.shape)))
    for t in range(st.T):
        trk.tracking(t, cams, [None] * len(cams), st.frame_boxes(t), st.frame_detections(t), "SVD")
        rep = [(tr.track_id, tr.poses3d[-1]["pose3d"]) for tr in trk.tracks
               if tr.time_since_update == 0 and tr.is_confirmed()]                 # ivclabpose.py:265-267
        k = g["count"][t]
        assert [r[0] for r in rep] == g["ids"][t, :k].tolist(), t
        for (tid, pose), ref in zip(rep, g["joints"][t, :k]):
            assert np.abs(pose - ref).max() < 5e-4
    ids = trk.tracks_ids                      # every id ever handed out, like the reference's never-pruned set
    assert set(g["ids"][g["ids"] >= 0].tolist()) <= ids and {tr.track_id for tr in trk.tracks} <= ids
    assert ids == set(range(len(ids)))
    assert set(trk.unmatched.keys()) == set(range(len(cams)))


def test_dropin_functions_match_reference_vectors():
    D = util.load_dropin()
    z = util.golden_functions()
    cams = camera.GetCameraParameters(dict(P=z["P"], K=z["K"], RT=z["RT"]))
    M, Cn, K, H = D.matching, D.construction, D.calculate, D.hypothesis
    assert np.array_equal(np.stack([c.F for c in cams]), z["cam_F"])
    px = dict(rtol=1e-9, atol=1e-6)       # pixels
    mm = dict(rtol=1e-3, atol=5e-4)       # metres: 0.5 mm / 1e-3 relative
    got = np.stack([c.projectPoints_parallel(z["proj_in"]) for c in cams])
    assert np.allclose(got, z["proj_out"], **px)
    sub = [cams[i] for i in z["eap_order"]]
    m, Dm = M.epipolar_affinity_parallel(sub, np.arange(len(sub)), z["eap_pose"], 17)
    assert np.allclose(m, z["eap_mean"], **px) and np.allclose(Dm, z["eap_D"], **px)
    m, Dm = M.epipolar_affinity(cams, z["ea_cam"], z["ea_pose"], 17)
    assert m.dtype == np.float32 and Dm.dtype == np.float32
    assert np.allclose(m, z["ea_mean"], rtol=1e-6, atol=1e-4) and np.allclose(Dm, z["ea_D"], rtol=1e-6, atol=1e-4)
    assert np.allclose(M.epipolar_distance(cams[1], z["pa"][1], cams[3], z["pb"][3]), z["ed_out"], **px)
    for b in range(len(z["gm_A"])):
        ml, bl, _ = M.Greedy_matching(sub, pose_mat=z["gm_pose"][b].reshape(-1, 1, 3), affinity_mat=z["gm_A"][b],
                                      next_pose=z["gm_next"][b])
        assert np.array_equal(bl[::2], z["gm_keep_update"][b]) and np.array_equal(ml, np.nonzero(bl[::2])[0])
        _, bl, _ = M.Greedy_matching(sub, affinity_mat=z["gm_A"][b].astype(np.float32), mode="init")
        assert np.array_equal(bl[::2], z["gm_keep_init"][b])
    keep = z["svd_keep"]
    jv = [[] for _ in sub]
    for j in range(17):
        jv[int(keep[j, ::2].sum()) - 1].append(j)
    Ts = list(z["svd_Ts"])
    assert np.allclose(Cn.SVD_pose_kernel_jf(sub, Ts, z["eap_pose"], 5, keep, jv, z["svd_next"]), z["svd_jf"], **mm)
    assert np.allclose(Cn.SVD_pose_kernel_parallel(sub, Ts, z["eap_pose"], 5), z["svd_parallel"], **mm)
    joints = [[z["eap_pose"][v, j] for v in range(len(sub))] for j in range(17)]
    remains = [[v for v in range(len(sub)) if keep[j, 2 * v]] for j in range(17)]
    old = np.array(Cn.SVD_pose_kernel(sub, Ts, joints, remains, 5, z["svd_next"]), dtype=np.float64)
    assert np.allclose(old, z["svd_old"], **mm)
    dirs = M.back_project_ray(cams[2].RK_INV, cams[2].position, z["ray_uv"])
    assert np.allclose(dirs, z["ray_dirs"], rtol=1e-12, atol=1e-12)
    assert np.allclose(K.line2point_distance_3D(cams[2].position, z["ray_dirs"], z["ray_X"]), z["ray_dist"], rtol=1e-9, atol=1e-9)
    hyp = H.Hypothesis(cams[0], z["pa"][0], 60)
    hyp.merge(cams[2], z["pa"][2])
    c1, v1 = hyp.calculate_cost(cams[3], z["pa"][3])
    c2, v2 = hyp.calculate_cost(cams[3], z["pb"][3])
    assert np.allclose([c1, c2], z["hyp_cost"], rtol=1e-9) and [v1, v2] == list(z["hyp_veto"])
    hyp.merge(cams[3], z["pa"][3])
    hyp.merge(cams[4], z["pa"][4])
    _, _, p3d, jvh, ok = hyp.get_3dpose_jf(30, 5)
    assert ok == bool(z["hyp_ok"]) and np.allclose(p3d, z["hyp_pose3d"], **mm)
    nvj = np.zeros(17, np.int32)
    for k, js in enumerate(jvh):
        nvj[js] = k + 1
    assert np.array_equal(nvj, z["hyp_views"])
    assert np.allclose([K.get_believe(z["pa"][0]), K.get_believe(z["pb"][1])], z["believe"])


def test_result_packing_equals_the_reference_facade():
    """f2 parity on the device: results.person_track_output on libpam's outputs (incl. the view-list output) equals
    the tuple the UNMODIFIED PersonTrack_Project3DPose returned (src/ivclabpose.py:259-287), frame by frame; the
    drop-in IterTrack objects carry the same joints_views / poses2d order."""
    import torch
    from pam_b200 import tracker
    g = util.golden_results()
    st = synth.make_stream(g["shape"], g["seq"], g["T"], **g["kw"])
    cams = camera.GetCameraParameters(st.rig)
    trk = tracker.SequenceTracker(cams, synth.tracker_params(st.shape), 1, max_detections=st.dets.shape[2], max_tracks=12,
                                  arm_joints=st.shape.arm_joints)
    out = trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda(), nviews=True,
                  assoc=True, vlist=True)
    assert trk.check().tolist() == [0]
    out = {k: v.cpu().numpy() for k, v in out.items()}
    util.check_results_against_facade(out, st.dets[None], g)
    host = trk.run_host(st.dets[None], st.counts[None], fresh=True, assoc=True, vlist=True)
    util.check_results_against_facade(host, st.dets[None], g)
    # the per-frame drop-in: what the facade's loop reads from tracker.tracks (ivclabpose.py:265-283)
    D = util.load_dropin()
    from types import SimpleNamespace
    dt = D.IterativeTracker.IterativeTracker(SimpleNamespace(**synth.tracker_params(st.shape)))
    for t in range(0, 40):
        dt.tracking(t, cams, [None] * len(cams), st.frame_boxes(t), st.frame_detections(t), "SVD")
        fr = g["frames"][t]
        rep = [tr for tr in dt.tracks if tr.time_since_update == 0 and tr.is_confirmed()]
        assert [tr.track_id for tr in rep] == fr["person3d_ids"], t
        for tr, jv, cam_ids, pids in zip(rep, fr["joints_views"], fr["camera_ids"], fr["person_ids"]):
            assert tr.poses3d[-1]["joints_views"] == jv, t
            assert [cid for cid, v in tr.poses2d.items() if v["time"] == t] == cam_ids, t
            assert len(tr.poses2d) == len(pids), t
