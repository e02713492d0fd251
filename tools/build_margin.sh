#!/bin/sh
# Diagnostic build of the library with the decision-margin counters:
#   sh tools/build_margin.sh   ->  csrc/libpam_margin.so   (use with PAM_LIBRARY=<that file>; pam_track_margins works)
# The product library (csrc/libpam.so, __graft_entry__.build()) is built WITHOUT them.
set -e
C="$(dirname "$0")/../part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc"
${NVCC:-/usr/local/cuda/bin/nvcc} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
    -DPAM_MARGIN -o "$C/libpam_margin.so" "$C/pam_lib.cu"
echo "built $C/libpam_margin.so"
