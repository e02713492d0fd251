"""PCP / MPJPE counters on the device: the nearest consumer of the
tracker's output and the quantity a multi-GPU run all-reduces.  Semantics of
``Evaluate3DPose_PCP`` (src/evalmodel.py:120-206)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .tracker import _check

BONE_GROUPS = {"Head": [8], "Torso": [9], "Upper arms": [5, 6], "Lower arms": [4, 7], "Upper legs": [1, 2],
               "Lower legs": [0, 3]}


def pcp_counters(trk, out, gt, gt_valid=None, frame_begin=0, frame_end=None, alpha=0.5, counters=None, mpjpe=None):
    """``trk``: SequenceTracker (its handle and joint count are used); ``out``: its output dict of
    CUDA tensors; ``gt`` (S,T,P,14,3) float64 CUDA tensor in Shelf/Campus joint order; ``gt_valid``
    (S,T,P) uint8.  Returns CUDA tensors ``counters (P,10,2) int64`` and ``mpjpe (2,) float64``
    (accumulated into the ones passed in, if any)."""
    import torch
    S, T, P = gt.shape[0], gt.shape[1], gt.shape[2]
    dev = gt.device
    assert gt.dtype == torch.float64 and gt.is_contiguous() and tuple(gt.shape[3:]) == (14, 3)
    if gt_valid is None:
        gt_valid = torch.ones((S, T, P), dtype=torch.uint8, device=dev)
    if counters is None:
        counters = torch.zeros((P, 10, 2), dtype=torch.int64, device=dev)
    if mpjpe is None:
        mpjpe = torch.zeros(2, dtype=torch.float64, device=dev)
    J = trk.cfg.num_joints
    st = torch.cuda.current_stream(dev).cuda_stream
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = trk.lib.pam_eval_pcp(trk.handle, p(out["count"]), p(out["joints"]), p(gt), p(gt_valid), S, T, P,
                              trk.out_rows, 1 if J == 17 else 0, frame_begin, T if frame_end is None else frame_end,
                              float(alpha), p(counters), p(mpjpe), C.c_void_p(st))
    _check(trk.lib, trk.handle, rc)
    return counters, mpjpe


def pcp_table(counters):
    """Per bone group / actor PCP and the total average (src/evalmodel.py:179-206)."""
    c = np.asarray(counters.cpu() if hasattr(counters, "cpu") else counters, dtype=np.float64)
    out = {}
    with np.errstate(invalid="ignore", divide="ignore"):
        for name, idx in BONE_GROUPS.items():
            out[name] = c[:, idx, 0].sum(1) / c[:, idx, 1].sum(1)
        out["Total"] = c[:, :, 0].sum(1) / c[:, :, 1].sum(1)
        out["total_avg"] = c[:, :, 0].sum() / c[:, :, 1].sum()
    return out


# ------------------------------------------------------------------------------------------------
# Panoptic: AP / recall / MPJPE (src/evalmodel.py:208-350)
# ------------------------------------------------------------------------------------------------
def panoptic_match(trk, out, gt_mm, gt_vis, n_gt):
    """Device matching step: ``gt_mm`` (S,T,G,14,3) float64 CUDA (millimetres), ``gt_vis`` (S,T,G,14)
    uint8, ``n_gt`` (S,T) int32 -> ``mpjpe`` (S,T,max_tracks) float64, ``gt_index`` (S,T,max_tracks) int32."""
    import torch
    S, T, G = gt_mm.shape[0], gt_mm.shape[1], gt_mm.shape[2]
    dev = gt_mm.device
    MT = trk.out_rows
    mp = torch.empty((S, T, MT), dtype=torch.float64, device=dev)
    gi = torch.empty((S, T, MT), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = trk.lib.pam_eval_panoptic_match(trk.handle, p(out["count"]), p(out["joints"]), p(gt_mm), p(gt_vis), p(n_gt),
                                         S, T, G, MT, p(mp), p(gi), C.c_void_p(st))
    _check(trk.lib, trk.handle, rc)
    return mp, gi


def eval_list_from_match(count, mpjpe, gt_index, n_gt, frames=None):
    """The reference's ``eval_list`` of one sequence: one ``(mpjpe, gt_id)`` per predicted pose, frames in
    order, ``gt_id`` numbered over all ground-truth bodies seen so far (evalmodel.py:304-320)."""
    count, mpjpe, gt_index, n_gt = (np.asarray(x.cpu() if hasattr(x, "cpu") else x) for x in (count, mpjpe, gt_index, n_gt))
    items, total = [], 0
    for t in (range(len(count)) if frames is None else frames):
        if n_gt[t] <= 0:
            continue
        for q in range(int(count[t])):
            items.append((float(mpjpe[t, q]), int(total + gt_index[t, q])))
        total += int(n_gt[t])
    return items, total


def panoptic_metrics(eval_list, total_gt, thresholds=tuple(range(25, 155, 25))):
    """-> (aps, recalls, mpjpe, recall@500) exactly as ``evaluate`` computes them (evalmodel.py:249-337)."""
    def ap_at(thr):
        n = len(eval_list)
        tp, fp, seen = np.zeros(n), np.zeros(n), set()
        for i, (m, g) in enumerate(eval_list):
            if m < thr and g not in seen:
                tp[i] = 1
                seen.add(g)
            else:
                fp[i] = 1
        tp, fp = np.cumsum(tp), np.cumsum(fp)
        recall = tp / (total_gt + 1e-5)
        precise = tp / (tp + fp + 1e-5)
        for k in range(n - 2, -1, -1):
            precise[k] = max(precise[k], precise[k + 1])
        precise = np.concatenate(([0], precise, [0]))
        recall = np.concatenate(([0], recall, [1]))
        idx = np.where(recall[1:] != recall[:-1])[0]
        return np.sum((recall[idx + 1] - recall[idx]) * precise[idx + 1]), recall[-2]
    aps, recs = zip(*[ap_at(t) for t in thresholds]) if thresholds else ((), ())
    seen, kept = set(), []
    for m, g in eval_list:
        if m < 500 and g not in seen:
            kept.append(m)
            seen.add(g)
    mp = float(np.mean(kept)) if kept else float("inf")
    rec500 = len({g for m, g in eval_list if m < 500}) / total_gt if total_gt else 0.0
    return list(aps), list(recs), mp, rec500
