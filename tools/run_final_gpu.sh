# Round-end validation on one B200: GPU tests, sanitizers on a tiny run, bench line, ncu launch list and one
# full capture of the tracker kernel (outputs under gpurun_out/; summarised into profiles/ by tools/summarise_ncu.py)
set -x
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 250 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/tiny_run.py shelf 40 3 2>&1 | tail -3
PAM_TRACK_THREADS=64 PAM_TRACK_MINBLOCKS=12 timeout 250 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/tiny_run.py shelf 40 3 2>&1 | tail -3
timeout 250 compute-sanitizer --tool memcheck python tools/tiny_run.py panoptic 30 2 2>&1 | tail -3
timeout 250 compute-sanitizer --tool initcheck python tools/tiny_run.py shelf 30 2 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 300 gpurun_out/bench_r01.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_track_sequences -c 1 -f -o gpurun_out/r01_track_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
