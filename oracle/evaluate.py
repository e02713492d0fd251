"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's PCP evaluation
(src/evalmodel.py:120-206 ``Evaluate3DPose_PCP``, src/eval/transformation.py:5-39 ``coco2shelf3D``,
src/eval/numeric.py:5-25 ``vectorize_distance``) on arrays instead of pickle / .mat files.

``tests/test_evaluate.py`` checks it against the unmodified reference function (through temporary
pickle + actorsGT.mat files) in the build container."""
import numpy as np

BONES = [[0, 1], [1, 2], [3, 4], [4, 5], [6, 7], [7, 8], [9, 10], [10, 11], [12, 13]]
BONE_GROUPS = {"Head": [8], "Torso": [9], "Upper arms": [5, 6], "Lower arms": [4, 7], "Upper legs": [1, 2],
               "Lower legs": [0, 3]}


def coco2shelf3D(coco_pose):
    """(3, 17) COCO order -> (14, 3) Shelf order (eval/transformation.py:5-39)."""
    coco = coco_pose.astype(float).T
    shelf = np.zeros((14, 3))
    shelf[0:12] += coco[np.array([16, 14, 12, 11, 13, 15, 10, 8, 6, 5, 7, 9])]
    shelf[12] = (shelf[8] + shelf[9]) / 2
    shelf[13] = shelf[12] + (coco[0] - shelf[12]) * np.array([0.78, 0.5, 1.5])
    shelf[12] = shelf[12] + (coco[0] - shelf[12]) * np.array([0.3, 0.4, 0.6])
    return shelf


def vectorize_distance(a, b):
    """eval/numeric.py:5-25."""
    N = a.shape[0]
    a = a.reshape(N, -1)
    dists = []
    for p in b:
        p = p.reshape(1, -1)
        remain = ~np.isnan(p)
        gt = a[remain].reshape(N, -1)
        gt2 = np.tile(np.sum(gt ** 2, axis=1).reshape(-1, 1), (1, 1))
        p = p[remain].reshape(1, -1)
        p2 = np.tile(np.sum(p ** 2, axis=1), (N, 1))
        d = gt2 + p2 - 2 * (gt @ p.T)
        dists.append(d / len(remain))
    return np.array(dists).reshape(1, -1)


def _is_right(ms, me, gs, ge, alpha=0.5):
    bone = np.linalg.norm(ge - gs)
    return ((np.linalg.norm(gs - ms) + np.linalg.norm(ge - me)) / 2) <= alpha * bone


def pcp_check(poses_per_frame, gt, gt_valid, frames, to_shelf=None):
    """``poses_per_frame[t]`` = (n, 14, 3) predictions in Shelf order (or (n, 3, 17) COCO with
    ``to_shelf=coco2shelf3D``); ``gt`` (T, P, 14, 3); ``gt_valid`` (T, P) -> check_result (T, P, 10)
    with +1 correct / -1 wrong / 0 not evaluated, exactly like evalmodel.py:146-177."""
    T, P = gt.shape[0], gt.shape[1]
    check = np.zeros((T, P, 10), dtype=np.int32)
    for t in frames:
        poses = poses_per_frame[t]
        for pid in range(P):
            if not gt_valid[t, pid]:
                continue
            if len(poses) == 0:
                check[t, pid, :] = -1
                continue
            model = np.stack([to_shelf(p) for p in poses]) if to_shelf else np.asarray(poses, dtype=float)
            g = gt[t, pid]
            dist = vectorize_distance(np.expand_dims(g, 0), model)
            m = model[np.argmin(dist[0])]
            for i, (s, e) in enumerate(BONES):
                check[t, pid, i] = 1 if _is_right(m[s], m[e], g[s], g[e]) else -1
            ghip, mhip = (g[2] + g[3]) / 2, (m[2] + m[3]) / 2
            check[t, pid, -1] = 1 if _is_right(mhip, m[12], ghip, g[12]) else -1
    return check


def counters_from_check(check):
    """(P, 10, 2): (correct, evaluated) per actor and part."""
    return np.stack([(check > 0).sum(0), np.abs(check).sum(0)], axis=-1).astype(np.int64)


def pcp_table(counters):
    """Per bone group and actor PCP + total, as evalmodel.py:179-206 tabulates it."""
    c = np.asarray(counters, dtype=np.float64)
    out = {}
    with np.errstate(invalid="ignore", divide="ignore"):
        for name, idx in BONE_GROUPS.items():
            out[name] = c[:, idx, 0].sum(1) / c[:, idx, 1].sum(1)
        out["Total"] = c[:, :, 0].sum(1) / c[:, :, 1].sum(1)
        out["total_avg"] = c[:, :, 0].sum() / c[:, :, 1].sum()
    return out


# ------------------------------------------------------------------------------------------------
# Panoptic evaluation (src/evalmodel.py:208-350), on arrays instead of json / pickle files
# ------------------------------------------------------------------------------------------------
def panoptic_eval_list(preds, gts):
    """``preds[t]`` = (n, 3, 17) predictions in metres (the pickle's arrays), ``gts[t]`` = dict with
    'joints_3d' (list of (14,3) mm) and 'joints_3d_vis' (list of (14,3) bool), frames in order
    (evalmodel.py:291-320) -> (eval_list of dicts, total_gt)."""
    eval_list, total_gt = [], 0
    for t, gt in gts.items():
        joints_3d, joints_3d_vis = gt["joints_3d"], gt["joints_3d_vis"]
        if len(joints_3d) == 0:
            continue
        for pose in preds[t].copy():
            pose = pose.T * 1000.
            pelvis = (pose[11] + pose[12]) / 2
            pose = pose[[0, 5, 7, 9, 11, 13, 15, 6, 8, 10, 12, 14, 16]]
            pose = np.insert(pose, 1 * 3, pelvis).reshape(-1, 3)
            mpjpes = []
            for (g, gvis) in zip(joints_3d, joints_3d_vis):
                vis = gvis[:, 0] > 0
                mpjpes.append(np.mean(np.sqrt(np.sum((pose[vis, 0:3] - g[vis]) ** 2, axis=-1))))
            eval_list.append({"mpjpe": float(np.min(mpjpes)), "gt_id": int(total_gt + np.argmin(mpjpes))})
        total_gt += len(joints_3d)
    return eval_list, total_gt


def panoptic_metrics(eval_list, total_gt):
    """AP@{25..150 mm}, recall, MPJPE, recall@500 (evalmodel.py:249-337)."""
    def to_ap(threshold):
        n = len(eval_list)
        tp, fp, gt_det = np.zeros(n), np.zeros(n), []
        for i, item in enumerate(eval_list):
            if item["mpjpe"] < threshold and item["gt_id"] not in gt_det:
                tp[i] = 1
                gt_det.append(item["gt_id"])
            else:
                fp[i] = 1
        tp, fp = np.cumsum(tp), np.cumsum(fp)
        recall = tp / (total_gt + 1e-5)
        precise = tp / (tp + fp + 1e-5)
        for k in range(n - 2, -1, -1):
            precise[k] = max(precise[k], precise[k + 1])
        precise = np.concatenate(([0], precise, [0]))
        recall = np.concatenate(([0], recall, [1]))
        index = np.where(recall[1:] != recall[:-1])[0]
        return np.sum((recall[index + 1] - recall[index]) * precise[index + 1]), recall[-2]
    aps, recs = [], []
    for t in np.arange(25, 155, 25):
        a, r = to_ap(t)
        aps.append(a)
        recs.append(r)
    gt_det, mp = [], []
    for item in eval_list:
        if item["mpjpe"] < 500 and item["gt_id"] not in gt_det:
            mp.append(item["mpjpe"])
            gt_det.append(item["gt_id"])
    mpjpe = np.mean(mp) if len(mp) > 0 else np.inf
    rec = len(np.unique([e["gt_id"] for e in eval_list if e["mpjpe"] < 500])) / total_gt
    return aps, recs, mpjpe, rec
