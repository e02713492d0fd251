"""Tiny tracker run for compute-sanitizer (development aid).
  [PAM_TRACK_SHAPE=...] [PAM_TINY_ALPHA_MULT=6] compute-sanitizer --tool racecheck python tools/tiny_run.py shelf 40 8"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pam_b200
from pam_b200 import synth, camera, tracker
shape = sys.argv[1] if len(sys.argv) > 1 else "shelf"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
S = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rig, dets, counts, gt, streams = synth.make_batch(shape, S, T, miss_prob=0.1, outlier_prob=0.05)
cams = camera.GetCameraParameters(rig)
params = synth.tracker_params(shape)
params["alpha2d"] *= float(os.environ.get("PAM_TINY_ALPHA_MULT", "1"))      # > 1: contested cameras in most frames (assignment paths)
trk = tracker.SequenceTracker(cams, params, S, max_detections=dets.shape[3],
                              max_tracks=8 if synth.SHAPES[shape].P <= 4 else 16, arm_joints=synth.SHAPES[shape].arm_joints)
out = trk.run(torch.from_numpy(dets).cuda(), torch.from_numpy(counts).cuda(), assoc=True, vlist=True, timing=True)
print("launch", trk.launch_info())
print("status", trk.check(strict=False), "reports", int(out["count"].sum()))
