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
timeout 300 python -m pytest tests/test_tracker_gpu.py tests/test_fuzz_hostemu.py tests/test_evaluate.py -m gpu -x -q 2>&1 | tail -3
