cd /root/repo
timeout 300 python -m pytest tests/test_stream_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gputest8.log
cat gpurun_out/r2_gputest8.log | tail -5
timeout 200 python - <<'PY' 2>&1 | tail -12
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch, json
import bench
print(json.dumps(bench.per_frame_api(torch.device('cuda:0'))))
from tests import util
D = util.load_dropin()
D.IterativeTracker.IterativeTracker.MAX_TRACKS = 8
D.IterativeTracker.IterativeTracker.MAX_DETECTIONS = 4
print(json.dumps(bench.per_frame_api(torch.device('cuda:0'))))
PY
