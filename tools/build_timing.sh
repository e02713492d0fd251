#!/bin/sh
# Development build with per-phase cycle accounting (printed by block 0 at kernel end):
#   sh tools/build_timing.sh && PAM_LIBRARY=$PWD/part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc/libpam_timing.so python tools/tiny_run.py shelf 800 1
cd "$(dirname "$0")/../part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc" && \
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DPAM_PHASE_TIMING -o libpam_timing.so pam_lib.cu
