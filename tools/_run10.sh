set -x
CFG="rgbbox:1000:1000:64,irreg:1000:1000:64,rgbbox:1000:1000:1,irreg:1000:1000:1,irreg:4000:4000:1,random:2000:2000:2:1000000"
RAY_B200_LIB=$PWD/raytracers_b200/_ab/libray_OLDMASK.so timeout 600 python tools/gpu_dev.py --tag ab_oldmask --reps 7 --configs $CFG --grid "warpqueue:" 2>&1 | grep config
timeout 600 python tools/gpu_dev.py --tag ab_newmask --reps 7 --configs $CFG --grid "warpqueue:" 2>&1 | grep config
RAY_B200_LIB=$PWD/raytracers_b200/_ab/libray_OLDMASK.so timeout 600 python tools/gpu_dev.py --tag ab_oldmask2 --reps 7 --configs $CFG --grid "warpqueue:" 2>&1 | grep config
python tools/trace_tail.py rgbbox irreg 2>&1 | grep "probes0x8" | cut -c1-260
