"""e2e (host buffers) rate for different chunk counts of the H2D / kernel / D2H pipeline."""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pam_b200
from pam_b200 import synth, camera, tracker
S, T = int(sys.argv[1]), 3200
rig, dets, counts, gt, streams = synth.make_batch("shelf", 8, T)
reps = (S + 7) // 8
h_dets = torch.empty((S, T, 5, 4, 14, 3), dtype=torch.float32, pin_memory=True)
h_counts = torch.empty((S, T, 5), dtype=torch.int32, pin_memory=True)
h_dets.numpy()[:] = np.tile(dets, (reps, 1, 1, 1, 1, 1))[:S]
h_counts.numpy()[:] = np.tile(counts, (reps, 1, 1))[:S]
cams = camera.GetCameraParameters(rig)
out = dict(count=torch.empty((S, T), dtype=torch.int32, pin_memory=True).numpy(),
           ids=torch.empty((S, T, 8), dtype=torch.int32, pin_memory=True).numpy(),
           joints=torch.empty((S, T, 8, 14, 3), dtype=torch.float32, pin_memory=True).numpy(), nviews=None, assoc=None)
# raw H2D bandwidth
d = torch.empty_like(h_dets, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h_dets, non_blocking=True); torch.cuda.synchronize()
print(f"raw 1-D H2D: {h_dets.numel()*4/(time.perf_counter()-t0)/1e9:.1f} GB/s")
del d
for nch in sys.argv[2].split(","):
    os.environ["PAM_HOST_CHUNKS"] = nch
    trk = tracker.SequenceTracker(cams, synth.tracker_params("shelf"), S, 4, 8, arm_joints=(6, 11))
    for _ in range(2):
        trk.run_host(h_dets.numpy(), h_counts.numpy(), fresh=True, nviews=False, out=out)
    t0 = time.perf_counter()
    for _ in range(3):
        trk.run_host(h_dets.numpy(), h_counts.numpy(), fresh=True, nviews=False, out=out)
    el = (time.perf_counter() - t0) / 3
    print(f"chunks={nch:>3s}: {el*1e3:7.1f} ms  {S*T/el/1e6:6.2f} M frames/s  H2D-equivalent {h_dets.numel()*4/el/1e9:.1f} GB/s")
    trk.close()
