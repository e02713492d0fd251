cd /root/repo
python -m pytest tests/test_tracker_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gputest3.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_track_sequences -s 1 -c 1 -f -o gpurun_out/r2_g1 python tools/one_shape.py 1:8:80 3552 3200 148 > gpurun_out/ncu_g1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_track_sequences -s 1 -c 1 -f -o gpurun_out/r2_g3 python tools/one_shape.py 3:1:80 1184 3200 148 > gpurun_out/ncu_g3.log 2>&1
tail -3 gpurun_out/r2_gputest3.log; tail -2 gpurun_out/ncu_g1.log gpurun_out/ncu_g3.log; ls -la gpurun_out/*.ncu-rep | tail -3
