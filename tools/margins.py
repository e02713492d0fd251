"""Decision margins of a run: the smallest distance any decision came to flipping.

  sh tools/build_margin.sh && PAM_LIBRARY=<csrc>/libpam_margin.so python tools/margins.py [--sequences 2368] [--frames 3200]

Runs the bench workload (Shelf-shaped sequences) through the -DPAM_MARGIN build of the library and prints, over all
sequences, the minimum of: |c| of "c > 0" (IterativeTracker.py:143), |A| of "A < 0" (matching.py:248), |ra - rb| / max of
the ray rule (matching.py:272), |believe - conf_threshold| (IterativeTracker.py:59), init-mode |A| (float32) and row-sum
difference (matching.py:287-294), |cost - 1| of the veto (hypothesis.py:66), smallest positive affinity.  FP64 rounding
differences between the device and numpy are ~1e-13 relative, so a decision can only flip when its margin is that small."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pam_b200  # noqa
from pam_b200 import camera, synth, tracker
from concurrent.futures import ThreadPoolExecutor

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="shelf")
ap.add_argument("--sequences", type=int, default=2368)
ap.add_argument("--frames", type=int, default=3200)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "margins.json"))
a = ap.parse_args()
sh = synth.SHAPES[a.shape]
rig = synth.make_rig(a.shape)
with ThreadPoolExecutor(min(32, os.cpu_count() or 8)) as ex:
    streams = list(ex.map(lambda s: synth.make_stream(a.shape, s, a.frames, rig=rig), range(a.sequences)))
dets = torch.from_numpy(np.stack([s.dets for s in streams])).cuda()
counts = torch.from_numpy(np.stack([s.counts for s in streams])).cuda()
trk = tracker.SequenceTracker(camera.GetCameraParameters(rig), synth.tracker_params(a.shape), a.sequences, max_detections=sh.P,
                              max_tracks=8 if sh.P <= 6 else 12, arm_joints=sh.arm_joints)
out = trk.run(dets, counts, nviews=False)
trk.check(strict=False)
m = trk.margins()
names = ["assoc |c|", "update |A|", "ray |ra-rb|/max", "|believe - conf_thr|", "init |A| (f32)", "init row-sum |s1-s2|",
         "veto |cost-1|", "smallest positive affinity"]
res = {"workload": f"{a.shape}: {a.sequences} sequences x {a.frames} frames (seeds 0..{a.sequences - 1})",
       "reports": int(out["count"].sum().item()),
       "margins": {n: (None if not np.isfinite(m[:, k].min()) else float(m[:, k].min())) for k, n in enumerate(names)},
       "sequence_of_minimum": {n: int(m[:, k].argmin()) for k, n in enumerate(names)},
       "note": "device-vs-numpy FP64 rounding differences are ~1e-13 relative; thresholds are O(1) quantities"}
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
print(json.dumps(res, indent=1))
