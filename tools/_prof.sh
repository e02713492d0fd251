cd /root/repo
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_track_sequences -c 1 -f -o gpurun_out/r02_track_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -1; tail -2 gpurun_out/bench_under_ncu_full.log | cut -c1-300
