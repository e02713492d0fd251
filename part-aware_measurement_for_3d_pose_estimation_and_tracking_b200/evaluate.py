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
                              trk.cfg.max_tracks, 1 if J == 17 else 0, frame_begin, T if frame_end is None else frame_end,
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
