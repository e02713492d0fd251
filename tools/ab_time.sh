# A/B kernel timing: csrc/libpam_base.so (built with an experiment macro off) against csrc/libpam.so
B=part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc/libpam_base.so
for rep in 1 2; do
  for S in 1,1184,1776; do
    echo "--- base"; PAM_LIBRARY=$B timeout 150 python tools/quick_time.py ${1:-shelf} 1000 $S 2>&1 | grep "S=" | cut -c1-100
    echo "--- new";  timeout 150 python tools/quick_time.py ${1:-shelf} 1000 $S 2>&1 | grep "S=" | cut -c1-100
  done
done
