set -x
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -2
for cfg in "64 8" "64 10" "64 12"; do set -- $cfg; PAM_TRACK_THREADS=$1 PAM_TRACK_MINBLOCKS=$2 timeout 120 python tools/quick_time.py shelf 1000 1184 2>&1 | tail -1 | cut -c1-100; done
timeout 120 python tools/quick_time.py shelf 1000 1184 2>&1 | tail -1 | cut -c1-100
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 600 gpurun_out/bench_r01.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_track_sequences -c 1 -f -o gpurun_out/r01_track_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
