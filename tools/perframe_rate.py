"""Rate of the per-frame drop-in API (IterativeTracker.tracking, one call per frame) and of the
bare C-ABI call underneath it."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util
from pam_b200 import synth, camera, tracker
from oracle.ref_loader import EasyDict
D = util.load_dropin()
st = synth.make_stream("shelf", 0, 1200)
cams = camera.GetCameraParameters(st.rig)
D.IterativeTracker.IterativeTracker.ARM_JOINTS = st.shape.arm_joints
trk = D.IterativeTracker.IterativeTracker(EasyDict(synth.tracker_params("shelf")))
inputs = [(st.frame_boxes(t), st.frame_detections(t)) for t in range(st.T)]
for t in range(200):
    trk.tracking(t, cams, [None] * 5, inputs[t][0], inputs[t][1], "SVD")
t0 = time.perf_counter()
for t in range(200, st.T):
    trk.tracking(t, cams, [None] * 5, inputs[t][0], inputs[t][1], "SVD")
el = (time.perf_counter() - t0) / (st.T - 200)
print(f"drop-in IterativeTracker.tracking(): {1/el:.0f} frames/s ({el*1e6:.1f} us/frame), reported ids {trk.last_ids.tolist()}")
# bare C-ABI call, one frame per call, numpy buffers prepared beforehand
raw = tracker.SequenceTracker(cams, synth.tracker_params("shelf"), 1, max_detections=4, max_tracks=8, arm_joints=st.shape.arm_joints)
out = None
fr = [(np.ascontiguousarray(st.dets[None, t:t + 1]), np.ascontiguousarray(st.counts[None, t:t + 1])) for t in range(st.T)]
for t in range(200):
    out = raw.run_host(fr[t][0], fr[t][1], frame0=t, fresh=(t == 0), nviews=False, out=out)
t0 = time.perf_counter()
for t in range(200, st.T):
    out = raw.run_host(fr[t][0], fr[t][1], frame0=t, nviews=False, out=out)
el = (time.perf_counter() - t0) / (st.T - 200)
print(f"pam_track_sequences_host, S=1 T=1 per call: {1/el:.0f} frames/s ({el*1e6:.1f} us/frame)")
