"""PCP evaluation: oracle restatement vs the unmodified reference function (build container only),
and the device counters vs the oracle (-m gpu)."""
import os
import pickle
import sys
import types

import numpy as np
import pytest

from tests import util
from oracle import evaluate as oeval, ref_loader
from pam_b200 import synth


def _noisy_predictions(rng, gt, drop=0.05):
    """Per frame a list of predicted poses: ground truth + noise, shuffled, sometimes missing."""
    T, P = gt.shape[:2]
    frames = []
    for t in range(T):
        poses = [gt[t, p] + rng.normal(0, 0.04, gt[t, p].shape) for p in range(P) if rng.random() > drop]
        if t % 17 == 5:
            poses = []
        rng.shuffle(poses)
        frames.append(np.array(poses).reshape(-1, gt.shape[2], 3))
    return frames


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the build container")
def test_oracle_pcp_matches_unmodified_reference(tmp_path):
    import scipy.io as scio
    rng = np.random.default_rng(4)
    T, P = 40, 3
    # COCO-17 ground truth -> the reference converts predictions with coco2shelf3D; GT is given in Shelf order
    st = synth.make_stream("shelf17", 9, T)
    gt17 = st.gt[:, :P]
    gt14 = np.stack([[oeval.coco2shelf3D(gt17[t, p].T) for p in range(P)] for t in range(T)])
    pred = _noisy_predictions(rng, gt17)
    valid = rng.random((T, P)) > 0.1
    # files in the reference's formats: {frame: (n, 3, J)} pickle and actorsGT.mat
    with open(tmp_path / "pred.pkl", "wb") as f:
        pickle.dump({t: np.transpose(pred[t], (0, 2, 1)) for t in range(T)}, f)
    actors = np.empty((1, P), dtype=object)
    for p in range(P):
        cells = np.empty((T, 1), dtype=object)
        for t in range(T):
            cells[t, 0] = gt14[t, p] if valid[t, p] else np.zeros((1, 0))
        actors[0, p] = cells
    scio.savemat(tmp_path / "actorsGT.mat", {"actor3D": actors})
    # import the reference's evalmodel.py with stubs for modules that are not installed
    src = os.path.join(ref_loader.REFERENCE_ROOT, "src")
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    try:
        for name in ("natsort", "motmetrics", "prettytable", "matplotlib", "matplotlib.pyplot", "dataset", "_init_path"):
            sys.modules[name] = types.ModuleType(name)

        class _PT:
            field_names = []
            def add_row(self, r): pass
            def __str__(self): return ""
        sys.modules["prettytable"].PrettyTable = _PT
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        for n in ("Testdatast", "GetConfig", "LoadFilenames", "LoadImages"):
            setattr(sys.modules["dataset"], n, None)
        if not hasattr(np, "float"):
            np.float = float
        sys.path[:0] = [src, os.path.join(src, "eval")]
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_evalmodel", os.path.join(src, "evalmodel.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ref_check, _ = mod.Evaluate3DPose_PCP([[0, T]], str(tmp_path / "pred.pkl"), gt_path=str(tmp_path), dataset_name="Shelf")
    finally:
        sys.path[:] = saved_path
        for k in ("natsort", "motmetrics", "prettytable", "matplotlib", "matplotlib.pyplot", "dataset", "_init_path",
                  "ref_evalmodel", "transformation", "numeric"):
            sys.modules.pop(k, None)
            if k in saved_mods:
                sys.modules[k] = saved_mods[k]
    mine = oeval.pcp_check([np.transpose(p, (0, 2, 1)) for p in pred], gt14, valid, range(T), to_shelf=oeval.coco2shelf3D)
    assert np.array_equal(ref_check, mine)
    assert (mine > 0).sum() > 300 and (mine < 0).sum() > 30


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["shelf", "shelf17"])
def test_device_pcp_counters_match_oracle(shape):
    import torch
    from pam_b200 import camera, evaluate, tracker
    S, T = 2, 80
    rig, dets, counts, gt, streams = synth.make_batch(shape, S, T, miss_prob=0.15, outlier_prob=0.1, noise_px=4.0)
    sh = synth.SHAPES[shape]
    trk = tracker.SequenceTracker(camera.GetCameraParameters(rig), synth.tracker_params(shape), S, dets.shape[3], 12,
                                  arm_joints=sh.arm_joints)
    out = trk.run(torch.from_numpy(dets).cuda(), torch.from_numpy(counts).cuda())
    trk.check()
    J = sh.J
    gt14 = gt if J == 14 else np.stack([[[oeval.coco2shelf3D(gt[s, t, p].T) for p in range(sh.P)] for t in range(T)]
                                        for s in range(S)])
    valid = (np.random.default_rng(0).random((S, T, sh.P)) > 0.1).astype(np.uint8)
    c, m = evaluate.pcp_counters(trk, out, torch.from_numpy(np.ascontiguousarray(gt14)).cuda(),
                                 torch.from_numpy(valid).cuda(), frame_begin=3)
    c = c.cpu().numpy()
    o = {k: v.cpu().numpy() for k, v in out.items() if v is not None}
    ref = np.zeros((sh.P, 10, 2), np.int64)
    for s in range(S):
        frames = [o["joints"][s, t, :o["count"][s, t]].astype(np.float64) for t in range(T)]
        if J == 17:
            frames = [np.transpose(f, (0, 2, 1)) for f in frames]
        chk = oeval.pcp_check(frames, gt14[s], valid[s], range(3, T), to_shelf=oeval.coco2shelf3D if J == 17 else None)
        ref += oeval.counters_from_check(chk)
    assert np.array_equal(c, ref)
    assert ref[:, :, 1].sum() > 1000 and 0.5 < ref[:, :, 0].sum() / ref[:, :, 1].sum() <= 1.0
    assert m.cpu().numpy()[1] > 0
    tab = evaluate.pcp_table(c)
    assert abs(tab["total_avg"] - oeval.pcp_table(ref)["total_avg"]) < 1e-12


# ------------------------------------------------------------------------------------------------
# Panoptic AP / recall / MPJPE
# ------------------------------------------------------------------------------------------------
def _panoptic_case(rng, T=30, P=3):
    st = synth.make_stream("shelf17", 12, T)
    gt17 = st.gt[:, :P]                                          # metres, COCO-17
    preds, gts = {}, {}
    for t in range(T):
        poses = [gt17[t, p] + rng.normal(0, 0.03, gt17[t, p].shape) for p in range(P) if rng.random() > 0.1]
        if t % 7 == 3:
            poses.append(gt17[t, 0] + 0.4)                       # a false positive
        rng.shuffle(poses)
        preds[t] = np.transpose(np.array(poses).reshape(-1, 17, 3), (0, 2, 1))
        j3d, vis = [], []
        for p in range(P):
            if rng.random() < 0.1:
                continue
            g = gt17[t, p] * 1000.0
            pelvis = (g[11] + g[12]) / 2
            g14 = np.vstack([g[0], pelvis, g[[5, 7, 9, 11, 13, 15, 6, 8, 10, 12, 14, 16]]])
            v = rng.random(14) > 0.15
            v[2] = True
            j3d.append(g14)
            vis.append(np.repeat(v.reshape(-1, 1), 3, axis=1))
        gts[t] = {"joints_3d": j3d, "joints_3d_vis": vis}
    return preds, gts


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the build container")
def test_oracle_panoptic_matches_unmodified_reference(tmp_path):
    """The reference's EvaluatePanoptic only prints its table; the rows it hands to PrettyTable and the
    MPJPE line are captured and compared (2 decimals, as printed)."""
    import contextlib, io, json
    rng = np.random.default_rng(8)
    preds, gts = _panoptic_case(rng)
    T = len(preds)
    # files in the reference's formats.  getGT keeps every 12th json; timestamps = frame * 12
    anno = tmp_path / "hdPose3d_stage1_coco19"
    anno.mkdir()
    Minv = np.linalg.inv(np.array([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]]))
    pk = {}
    for t in range(T):
        bodies = []
        for g14, v in zip(gts[t]["joints_3d"], gts[t]["joints_3d_vis"]):
            j19 = np.zeros((19, 4))
            j19[1:15, :3] = (g14 / 10.0) @ Minv                  # undo the x10 and the axis change of getGT
            j19[1:15, 3] = np.where(v[:, 0], 1.0, 0.0)
            bodies.append({"joints19": j19.reshape(-1).tolist()})
        for k in range(12):                                      # 11 filler files that getGT skips
            with open(anno / f"body3DScene_{t * 12 + k:08d}.json", "w") as f:
                json.dump({"bodies": bodies if k == 0 else []}, f)
        pk[t * 12] = preds[t]
    with open(tmp_path / "pred.pkl", "wb") as f:
        pickle.dump(pk, f)
    src = os.path.join(ref_loader.REFERENCE_ROOT, "src")
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    rows = []
    try:
        for name in ("natsort", "motmetrics", "prettytable", "matplotlib", "matplotlib.pyplot", "dataset", "_init_path"):
            sys.modules[name] = types.ModuleType(name)

        class _PT:
            field_names = []
            def add_row(self, r): rows.append(list(r))
            def __str__(self): return ""
        sys.modules["prettytable"].PrettyTable = _PT
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        for n in ("Testdatast", "GetConfig", "LoadFilenames", "LoadImages"):
            setattr(sys.modules["dataset"], n, None)
        if not hasattr(np, "float"):
            np.float = float
        sys.path[:0] = [src, os.path.join(src, "eval")]
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_evalmodel", os.path.join(src, "evalmodel.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            mod.EvaluatePanoptic([[0, T]], str(tmp_path / "pred.pkl"), data_root=str(tmp_path))
    finally:
        sys.path[:] = saved_path
        for k in ("natsort", "motmetrics", "prettytable", "matplotlib", "matplotlib.pyplot", "dataset", "_init_path",
                  "ref_evalmodel", "transformation", "numeric"):
            sys.modules.pop(k, None)
            if k in saved_mods:
                sys.modules[k] = saved_mods[k]
    ev, total = oeval.panoptic_eval_list(preds, gts)
    aps, recs, mpjpe, _ = oeval.panoptic_metrics(ev, total)
    assert rows[0] == ["AP"] + [f"{a * 100:.2f}" for a in aps]
    assert rows[1] == ["Recall"] + [f"{r * 100:.2f}" for r in recs]
    assert f"MPJPE: {mpjpe:.2f}mm" in buf.getvalue()
    assert 0.3 < aps[-1] <= 1.0 and total > 50


@pytest.mark.gpu
def test_device_panoptic_matching_and_metrics_match_oracle():
    import torch
    from pam_b200 import camera, evaluate, tracker
    rng = np.random.default_rng(5)
    S, T = 1, 60
    rig, dets, counts, gt, streams = synth.make_batch("shelf17", S, T, miss_prob=0.1, outlier_prob=0.05, noise_px=3.0)
    sh = synth.SHAPES["shelf17"]
    trk = tracker.SequenceTracker(camera.GetCameraParameters(rig), synth.tracker_params("shelf17"), S, dets.shape[3], 12,
                                  arm_joints=sh.arm_joints)
    out = trk.run(torch.from_numpy(dets).cuda(), torch.from_numpy(counts).cuda())
    trk.check()
    o = {k: v.cpu().numpy() for k, v in out.items() if v is not None}
    # ground truth in the evaluator's format (mm, 14 joints, visibility), some bodies / joints hidden
    G = sh.P
    gt_mm = np.zeros((S, T, G, 14, 3)); vis = np.zeros((S, T, G, 14), np.uint8); n_gt = np.zeros((S, T), np.int32)
    gts, preds = {}, {}
    for t in range(T):
        j3d, jv = [], []
        for p in range(sh.P):
            if rng.random() < 0.1:
                continue
            g = gt[0, t, p] * 1000.0
            g14 = np.vstack([g[0], (g[11] + g[12]) / 2, g[[5, 7, 9, 11, 13, 15, 6, 8, 10, 12, 14, 16]]])
            v = rng.random(14) > 0.15
            k = len(j3d)
            gt_mm[0, t, k], vis[0, t, k] = g14, v
            j3d.append(g14); jv.append(np.repeat(v.reshape(-1, 1), 3, axis=1))
        n_gt[0, t] = len(j3d)
        gts[t] = {"joints_3d": j3d, "joints_3d_vis": jv}
        preds[t] = np.transpose(o["joints"][0, t, :o["count"][0, t]].astype(np.float64), (0, 2, 1))
    mp, gi = evaluate.panoptic_match(trk, out, torch.from_numpy(gt_mm).cuda(), torch.from_numpy(vis).cuda(),
                                     torch.from_numpy(n_gt).cuda())
    items, total = evaluate.eval_list_from_match(o["count"][0], mp[0], gi[0], n_gt[0])
    ref_list, ref_total = oeval.panoptic_eval_list(preds, gts)
    assert total == ref_total and len(items) == len(ref_list) > 100
    assert [g for _, g in items] == [e["gt_id"] for e in ref_list]
    assert np.allclose([m for m, _ in items], [e["mpjpe"] for e in ref_list], rtol=1e-12, atol=1e-9)
    a, r, m, r5 = evaluate.panoptic_metrics(items, total)
    ra, rr, rm, rr5 = oeval.panoptic_metrics(ref_list, ref_total)
    assert np.allclose(a, ra) and np.allclose(r, rr) and abs(m - rm) < 1e-9 and abs(r5 - rr5) < 1e-12
