cd /root/repo
timeout 300 python -m pytest tests/test_tracker_gpu.py -m gpu -x -q -k "max_report" 2>&1 | grep -E "^E|assert|Error" | head -12
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/bench13.json 2> gpurun_out/bench13.err; tail -c 300 gpurun_out/bench13.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench13.json').read().strip().splitlines()[-1])
print("value", d["value"]/1e6, "kernel_ms", d["roofline"]["kernel_ms"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"]/1e6)
PY
