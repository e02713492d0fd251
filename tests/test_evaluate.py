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
