"""Per-frame device cycles of every sequence (the kernel's timing output) -> how much of a convoy's frame is waiting
for its slowest sequence.   python tools/convoy_stats.py [shape] [S] [T]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pam_b200  # noqa
from pam_b200 import camera, synth, tracker
from concurrent.futures import ThreadPoolExecutor
name = sys.argv[1] if len(sys.argv) > 1 else "shelf"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
T = int(sys.argv[3]) if len(sys.argv) > 3 else 600
sh = synth.SHAPES[name]
rig = synth.make_rig(name)
with ThreadPoolExecutor(16) as ex:
    streams = list(ex.map(lambda s: synth.make_stream(name, s, T, rig=rig), range(min(296, S))))
du = torch.from_numpy(np.stack([s.dets for s in streams])).cuda()
cu = torch.from_numpy(np.stack([s.counts for s in streams])).cuda()
reps = (S + len(streams) - 1) // len(streams)
dets, counts = du.repeat(reps, 1, 1, 1, 1, 1)[:S].contiguous(), cu.repeat(reps, 1, 1)[:S].contiguous()
trk = tracker.SequenceTracker(camera.GetCameraParameters(rig), synth.tracker_params(name), S, max_detections=dets.shape[3],
                              max_tracks=8 if sh.P <= 6 else 12, arm_joints=sh.arm_joints)
out = trk.run(dets, counts, timing=True)
torch.cuda.synchronize()
info = trk.launch_info()
Q = info["sequences_per_cta"]
tm = out["timing"].cpu().numpy().astype(np.float64)[:, 50:]          # [S][T][4]: association, update, initialisation, frame
print(info)
n = S // Q * Q
for k, lab in enumerate(("association", "update", "initialisation", "whole frame")):
    x = tm[:n, :, k].reshape(n // Q, Q, -1)
    mean, mx = x.mean(), x.max(axis=1).mean()
    print(f"{lab:15s} mean {mean:9.0f} cycles   mean over convoys of the slowest {mx:9.0f}   ratio {mx / max(mean, 1):.3f}   "
          f"p50 {np.percentile(x, 50):.0f} p90 {np.percentile(x, 90):.0f} p99 {np.percentile(x, 99):.0f} max {x.max():.0f}")
w = tm[:n, :, 3].reshape(n // Q, Q, -1)
arg = w.argmax(axis=1)                                                # which sequence is the slowest
slow = np.take_along_axis(tm[:n].reshape(n // Q, Q, -1, 4), arg[:, None, :, None], axis=1)[:, 0]
print("slowest sequence of a convoy frame: association %.0f update %.0f initialisation %.0f of %.0f cycles" %
      (slow[..., 0].mean(), slow[..., 1].mean(), slow[..., 2].mean(), slow[..., 3].mean()))
print("frames of a convoy in which some sequence runs the initialisation for > 2000 cycles: %.1f %%" %
      (100.0 * (tm[:n, :, 2].reshape(n // Q, Q, -1).max(axis=1) > 2000).mean()))
