set -x
timeout 900 python tools/gpu_dev.py --tag k2_sweep2 --reps 5 --configs rgbbox:1000:1000:64,rgbbox:1000:1000:1 \
  --grid "warpqueue:;warpqueue:wq_k=2,wq_warps=22|24|26,wq_ncap=256|320|384|448" 2>&1 | grep config
