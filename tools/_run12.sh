cd /root/repo
python tools/sweep_shapes.py --out gpurun_out/sweep12.json 1:8:128/c@2368 1:8:128:0/c@2368 1:12:80/c@3552 1:8:80/c@3552 3:1:80@1184 auto@1 auto@2368 > gpurun_out/sweep12.log 2>&1
grep -h -e mfps -e rror gpurun_out/sweep12.log | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    i = r['info']
    print(r['config'], round(r['ms'],1), 'ms', round(r['mfps'],1), 'M', 'G%d Q%d regs%d ctas/SM %d smem %d nbuf %d' % (i['warps_per_sequence'], i['sequences_per_cta'], i['registers_per_thread'], i['ctas_per_sm'], i['smem_per_cta'], i['detection_buffers']), r['identical_to_first'], r['clocks'])
"
timeout 600 python -m pytest tests -m gpu -x -q -k "not full_length_oracle" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench12.json 2> gpurun_out/bench12.err; tail -c 400 gpurun_out/bench12.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench12.json').read().strip().splitlines()[-1])
print("value", d["value"]/1e6, "kernel_ms", d["roofline"]["kernel_ms"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"]/1e6, "parity", d["parity_sample"], "per_frame", d["per_frame_api"]["value"], d["per_frame_api"]["stream_mode"]["value"])
PY
