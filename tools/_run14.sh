cd /root/repo
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_track_sequences -s 1 -c 1 -f -o gpurun_out/r02_track_full python tools/one_shape.py 1:8:128 2368 3200 148 > gpurun_out/ncu14.log 2>&1
tail -2 gpurun_out/ncu14.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu14.log 2>&1
tail -c 200 gpurun_out/bench_under_ncu14.log
for tool in racecheck memcheck; do
  PAM_TRACK_SHAPE=1:4:128 timeout 300 compute-sanitizer --tool $tool python tools/one_shape.py 1:4:128 8 40 8 2>&1 | tail -2
done
timeout 300 compute-sanitizer --tool racecheck python - <<'PY' 2>&1 | tail -3
import sys; sys.path.insert(0, '.')
import numpy as np
import pam_b200
from pam_b200 import camera, synth, tracker
st = synth.make_stream("shelf", 3, 40, miss_prob=0.1, outlier_prob=0.05, enter_stagger=5)
trk = tracker.SequenceTracker(camera.GetCameraParameters(st.rig), synth.tracker_params("shelf"), 1, max_detections=4, max_tracks=8, arm_joints=st.shape.arm_joints)
fs = trk.open_stream(True)
for t in range(st.T):
    fs.set_frame(st.frame_detections(t)); fs.step(t)
print("stream mode under racecheck: reports", int(fs.count[0]))
trk.close()
PY
ls -la gpurun_out/r02_track_full.ncu-rep
