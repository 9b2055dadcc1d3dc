set -x
timeout 900 python tools/gpu_dev.py --tag learn_order2 --reps 7 --configs rgbbox:1000:1000:1,irreg:1000:1000:1,irreg:4000:4000:1,rgbbox:1000:1000:64,irreg:1000:1000:64 \
  --grid "warpqueue:learn_order=0;warpqueue:learn_order=1,long_path=4;warpqueue:learn_order=1,long_path=-4;warpqueue:learn_order=1,long_path=-2;warpqueue:learn_order=1,long_path=-6" 2>&1 | grep config
