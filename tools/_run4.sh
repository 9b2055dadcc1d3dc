set -x
CFG="irreg:1000:1000:64,irreg:1000:1000:1,irreg:4000:4000:1,rgbbox:1000:1000:64,random:2000:2000:2:1000000"
for v in NO_LDG256 NO_TAILPF; do
RAY_B200_LIB=$PWD/raytracers_b200/_ab/libray_$v.so timeout 600 python tools/gpu_dev.py --tag ab_$v --reps 7 --configs $CFG --grid "warpqueue:" 2>&1 | grep config
done
timeout 600 python tools/gpu_dev.py --tag ab_new2 --reps 7 --configs $CFG --grid "warpqueue:" 2>&1 | grep config
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "peer_frame" 2>&1 | tail -n 15
