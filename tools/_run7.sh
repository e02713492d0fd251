cd /root/repo
python -m pytest tests -m gpu -x -q -k "not full_length_oracle" 2>&1 | tail -6 > gpurun_out/r2_gputest7.log
tail -3 gpurun_out/r2_gputest7.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
tail -c 6000 gpurun_out/bench7.json; tail -5 gpurun_out/bench7.err
