cd /root/repo
C=part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gputest10.log ) &
TESTPID=$!
PAM_LIBRARY=$C/libpam_margin.so timeout 600 python tools/margins.py --out gpurun_out/r02_margins.json > gpurun_out/margins.log 2>&1
tail -22 gpurun_out/margins.log
wait $TESTPID
tail -4 gpurun_out/r2_gputest10.log
timeout 900 python tools/gpu_fuzz.py --seeds 0 1200 --large 0 150 --log gpurun_out/r02_fuzz_gpu.log 2>&1 | tail -3
