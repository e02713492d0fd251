"""Rate of the per-frame drop-in API (IterativeTracker.tracking, one call per frame)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import util
from pam_b200 import synth, camera
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
el = time.perf_counter() - t0
print(f"drop-in IterativeTracker.tracking(): {1000/ (el/1000*1000/ (st.T-200)) :.0f} frames/s ({el/(st.T-200)*1e6:.1f} us/frame), reported ids {trk.last_ids.tolist()}")
