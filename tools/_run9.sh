cd /root/repo
timeout 300 python -m pytest tests/test_stream_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 200 python - <<'PY' 2>&1 | tail -12
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch, json
import pam_b200
from pam_b200 import camera, synth, tracker
st = synth.make_stream("shelf", 0, 2200)
cams = camera.GetCameraParameters(st.rig)
for D, MT in ((4, 8), (8, 12)):
    trk = tracker.SequenceTracker(cams, synth.tracker_params("shelf"), 1, max_detections=D, max_tracks=MT, arm_joints=st.shape.arm_joints)
    fs = trk.open_stream(True)
    inputs = [st.frame_detections(t) for t in range(st.T)]
    for t in range(200):
        fs.set_frame(inputs[t]); fs.step(t)
    acc = np.zeros(8); tset = tstep = 0.0
    for t in range(200, st.T):
        a = time.perf_counter(); fs.set_frame(inputs[t]); b = time.perf_counter(); fs.step(t); c = time.perf_counter()
        tset += b - a; tstep += c - b; acc += fs.timing
    n = st.T - 200; khz = trk.sm_clock_khz()
    tm = acc / n
    print(f"D={D} MT={MT}: set_frame {tset/n*1e6:.1f} us, step {tstep/n*1e6:.1f} us; device: input {tm[4]/khz*1e3:.1f} us, frame {tm[5]/khz*1e3:.1f} us (frame_step {tm[3]/khz*1e3:.1f}), "
          f"writeback {(int(fs.timing[6]) & 0xffff)/khz*1e3:.1f} us, persist {(int(fs.timing[6]) >> 16)/khz*1e3:.1f} us")
    trk.close()
import bench
print(json.dumps(bench.per_frame_api(torch.device('cuda:0')))[:900])
PY
