# A/B kernel timing of development builds: csrc/libpam_<name>.so ... against csrc/libpam.so ("new")
#   sh tools/ab_time.sh shelf base pair2 ...
C=part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc
shape=${1:-shelf}; shift
for rep in 1 2; do
  for v in "$@" new; do
    if [ $v = new ]; then lib=$C/libpam.so; else lib=$C/libpam_$v.so; fi
    echo "--- $v"; PAM_LIBRARY=$lib timeout 150 python tools/quick_time.py $shape 1000 1,1184,1776 2>&1 | grep "S=" | cut -c1-100
  done
done
