"""Result packing (section 8f-2) on synthetic output tensors; CPU only."""
import pickle

import numpy as np

from pam_b200 import results


def _fake_out():
    rng = np.random.default_rng(0)
    S, T, MT, J, V, D = 1, 3, 4, 14, 3, 2
    out = dict(count=np.array([[0, 2, 1]], np.int32), ids=np.zeros((S, T, MT), np.int32),
               joints=rng.normal(size=(S, T, MT, J, 3)).astype(np.float32), nviews=np.full((S, T, MT, J), 3, np.uint8),
               assoc=np.full((S, T, V, D), -1, np.int32))
    out["ids"][0, 1, :2] = [5, 7]
    out["ids"][0, 2, :1] = [7]
    out["nviews"][0, 1, 0, 4] = 2
    out["assoc"][0, 1, 0, 1] = 5
    out["assoc"][0, 1, 2, 0] = 5
    out["assoc"][0, 1, 1, 0] = 7
    dets = rng.normal(size=(S, T, V, D, J, 3)).astype(np.float32)
    return out, dets


def test_person_track_output_and_pickle(tmp_path):
    out, dets = _fake_out()
    cams, pts, pids, pts3d, views, ids = results.person_track_output(out, 0, 1, dets, n_views=3)
    assert ids.tolist() == [5, 7] and pts3d.shape == (2, 3, 14)
    assert np.array_equal(pts3d[1], out["joints"][0, 1, 1].T.astype(np.float64))
    assert views[0][1] == [4] and 4 not in views[0][2] and len(views[0][2]) == 13
    assert list(cams[0]) == [0, 2] and pids[0] == [5, 5] and list(cams[1]) == [1]
    assert np.array_equal(pts[0][1], dets[0, 1, 2, 0].astype(np.float64))
    mp = results.multi_poses3d(out, 0, frame_ids=[10, 11, 12])
    assert mp[10].shape == (0, 3, 14) and mp[11].shape == (2, 3, 14) and mp[12].shape == (1, 3, 14)
    results.write_3d_result(mp, str(tmp_path / "res" / "pred.pkl"))
    back = pickle.load(open(tmp_path / "res" / "pred.pkl", "rb"))
    assert sorted(back) == [10, 11, 12] and np.array_equal(back[11], mp[11])
    ann = [{"timestamp": "0001", "cid": 0, "pid": 5, "pose": dets[0, 1, 0, 1][:, :2], "scores": dets[0, 1, 0, 1][:, 2]}]
    results.write_2d_result((640, 480), ann, save_dir=str(tmp_path / "trk"))
    import json
    j = json.load(open(tmp_path / "trk" / "Camera0.json"))
    assert j["image_wh"] == [480, 640] and len(j["frames"]) == 1


def test_person_track_output_equals_the_reference_facade_hostemu():
    """f2 parity: the tuple packed from the kernel's outputs (kernel source on the host)
    equals what the UNMODIFIED PersonTrack_Project3DPose returned on top of the unmodified tracker, frame by frame:
    camera ids in dict-insertion order, person ids per view-dict entry, joints_views buckets, 3-D poses."""
    from tests import util
    from pam_b200 import synth
    g = util.golden_results()
    st = synth.make_stream(g["shape"], g["seq"], g["T"], **g["kw"])
    out = util.run_hostemu([st], util.stream_config(st, max_tracks=12))
    assert out["status"].tolist() == [0]
    util.check_results_against_facade(out, st.dets[None], g)
