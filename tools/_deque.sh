set -x
CFG="rgbbox:1000:1000:64,irreg:1000:1000:64,rgbbox:1000:1000:1,irreg:1000:1000:1,rgbbox:2000:2000:16,random:2000:2000:2:1000000,irreg:4000:4000:1"
RAY_B200_LIB=$PWD/raytracers_b200/_ab/libray_lifo.so timeout 600 python tools/gpu_dev.py --tag lifo --reps 5 --configs $CFG --grid "warpqueue:" 2>&1 | grep config
timeout 900 python tools/gpu_dev.py --tag deque --reps 5 --configs $CFG --grid "warpqueue:wq_low=-1|0|64|96|160" 2>&1 | grep config
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_deque.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/r2_test_deque.log
