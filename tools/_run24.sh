set -x
timeout 900 python tools/gpu_dev.py --tag slot80 --reps 5 --configs rgbbox:1000:1000:64,irreg:1000:1000:64,rgbbox:2000:2000:16 \
  --grid "warpqueue:;warpqueue:wq_warps=24|26|28|30,wq_k=2,wq_ncap=256" 2>&1 | grep config
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -n 4
