cd /root/repo
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2_gputest17.log 2>&1
cat gpurun_out/r2_gputest17.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1200 python bench.py > gpurun_out/bench17.json 2> gpurun_out/bench17.err ) 2>&1 | tail -3
tail -c 300 gpurun_out/bench17.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench17.json').read().strip().splitlines()[-1])
print("value", d["value"]/1e6, "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "secondary", d["roofline"]["secondary"], "traffic", d["roofline"]["traffic"])
print("e2e", d["e2e"]["value"]/1e6, "parity", d["parity_sample"]["ids_counts_nviews_assoc"], d["parity_sample"]["max_dx_m"], "per_frame", d["per_frame_api"]["value"], d["per_frame_api"]["stream_mode"]["value"], d["per_frame_api"]["launch_per_call"]["value"])
for k,v in d["other_configs"].items(): print(k, v.get("value"), v.get("error"))
PY
