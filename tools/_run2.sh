cd /root/repo
python -m pytest tests -m gpu -x -q -k "not full_length_oracle" 2>&1 | tail -15 > gpurun_out/r2_gputest2.log
python tools/sweep_shapes.py --out gpurun_out/sweep2.json \
  3:1:80@1184 3:1:80:1@1184 2:2:80@1184 1:8:128@1184 1:4:128@1184 \
  3:2:80@2368 2:2:80:1@2368 1:8:128@2368 1:8:80@2368 \
  1:8:80@3552 1:8:64@3552 1:7:72@3552 2:4:64@3552 1:8:128@3552 \
  1:7:72@4144 1:8:80@4144 1:8:64@4736 auto@1 auto@1184 auto@4144 > gpurun_out/sweep2.log 2>&1
tail -30 gpurun_out/sweep2.log | cut -c1-400
