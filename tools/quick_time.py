"""Quick kernel-only timing of pam_track_sequences for a few batch sizes (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pam_b200
from pam_b200 import synth, camera, tracker

shape = sys.argv[1] if len(sys.argv) > 1 else "shelf"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3200
sizes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 148, 592, 1184]
K = 8
t0 = time.time()
rig, dets, counts, gt, streams = synth.make_batch(shape, K, T)
print(f"generated {K} x {T} frames in {time.time()-t0:.1f}s", flush=True)
cams = camera.GetCameraParameters(rig)
sh = synth.SHAPES[shape]
for S in sizes:
    reps = (S + K - 1) // K
    d = torch.from_numpy(np.tile(dets, (reps, 1, 1, 1, 1, 1))[:S]).cuda()
    c = torch.from_numpy(np.tile(counts, (reps, 1, 1))[:S]).cuda()
    trk = tracker.SequenceTracker(cams, synth.tracker_params(shape), S, max_detections=dets.shape[3], max_tracks=8,
                                  arm_joints=sh.arm_joints)
    out = trk.alloc_outputs(T, nviews=False, assoc=False)
    best = 1e9
    for it in range(3):
        trk.restart()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        trk.run(d, c, out=out, frame0=0)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    trk.check()
    nrep = int(out["count"].sum().item())
    print(f"S={S:5d} T={T} threads={os.environ.get('PAM_TRACK_THREADS','auto')}: {best:9.2f} ms  {S*T/best*1e3:12.0f} frames/s "
          f"({best/T*1e3:7.2f} us/frame-step)  reports={nrep}", flush=True)
    del trk, d, c, out
