"""Run ONE launch shape of the tracker kernel a few times (target of ncu captures).
  python tools/one_shape.py <G:Q:regs[:lean]|auto> <S> [frames] [unique] [shape]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
shp, S = sys.argv[1], int(sys.argv[2])
T = int(sys.argv[3]) if len(sys.argv) > 3 else 3200
U = int(sys.argv[4]) if len(sys.argv) > 4 else 148
name = sys.argv[5] if len(sys.argv) > 5 else "shelf"
if shp != "auto":
    os.environ["PAM_TRACK_SHAPE"] = shp
import pam_b200  # noqa
from pam_b200 import camera, synth, tracker
from concurrent.futures import ThreadPoolExecutor
sh = synth.SHAPES[name]
rig = synth.make_rig(name)
with ThreadPoolExecutor(16) as ex:
    streams = list(ex.map(lambda s: synth.make_stream(name, s, T, rig=rig), range(min(U, S))))
du = torch.from_numpy(np.stack([s.dets for s in streams])).cuda()
cu = torch.from_numpy(np.stack([s.counts for s in streams])).cuda()
reps = (S + len(streams) - 1) // len(streams)
dets, counts = du.repeat(reps, 1, 1, 1, 1, 1)[:S].contiguous(), cu.repeat(reps, 1, 1)[:S].contiguous()
trk = tracker.SequenceTracker(camera.GetCameraParameters(rig), synth.tracker_params(name), S, max_detections=dets.shape[3],
                              max_tracks=8 if sh.P <= 6 else 12, arm_joints=sh.arm_joints)
out = trk.alloc_outputs(T, nviews=False, assoc=False)
for it in range(3):
    trk.restart()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); trk.run(dets, counts, out=out, frame0=0); e1.record(); torch.cuda.synchronize()
    print(f"{shp}@{S}: {e0.elapsed_time(e1):.2f} ms  {S * T / e0.elapsed_time(e1) / 1e3:.1f} M frames/s", trk.launch_info(), flush=True)
trk.check(strict=False)
