set -x
CFG="rgbbox:1000:1000:64,irreg:1000:1000:64,rgbbox:1000:1000:1,irreg:1000:1000:1,irreg:4000:4000:1,random:2000:2000:2:1000000"
RAY_B200_LIB=$PWD/raytracers_b200/_ab/libray_r1.so timeout 600 python tools/gpu_dev.py --tag ab_r1 --reps 7 --configs $CFG --grid "warpqueue:" 2>&1 | grep config
timeout 600 python tools/gpu_dev.py --tag ab_new --reps 7 --configs $CFG --grid "warpqueue:;lanewalk:lw_slots=48|64,lw_idle_min=16" 2>&1 | grep config
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_all2.log 2>&1; echo "all tests rc=$?"; tail -n 5 gpurun_out/r2_test_all2.log
