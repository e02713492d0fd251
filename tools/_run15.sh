cd /root/repo
C=part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc
for lib in libpam.so libpam_u2.so; do
  echo "== $lib"
  PAM_LIBRARY=$C/$lib python tools/sweep_shapes.py --out gpurun_out/sweep15_$lib.json 1:8:128/c@2368 1:10:96/c@2960 1:12:80/c@3552 2>&1 | grep mfps | python -c "
import sys, json
for ln in sys.stdin:
    r = json.loads(ln); i = r['info']
    print(r['config'], round(r['ms'],1), 'ms', round(r['mfps'],1), 'M', 'regs%d ctas/SM %d' % (i['registers_per_thread'], i['ctas_per_sm']), r['identical_to_first'])
"
done
timeout 300 python -m pytest tests/test_evaluate.py tests/test_tracker_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/bench15.json 2> gpurun_out/bench15.err; tail -c 300 gpurun_out/bench15.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench15.json').read().strip().splitlines()[-1])
print("value", d["value"]/1e6, "kernel_ms", d["roofline"]["kernel_ms"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"]/1e6)
PY
