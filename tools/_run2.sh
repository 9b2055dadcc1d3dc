set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lanewalk or small or golden" > gpurun_out/r2_test_lw.log 2>&1; echo "lanewalk tests rc=$?"
tail -n 15 gpurun_out/r2_test_lw.log
timeout 600 python tools/gpu_dev.py --tag lw1 --reps 5 --configs rgbbox:1000:1000:64,irreg:1000:1000:64,rgbbox:1000:1000:1,irreg:1000:1000:1 \
  --grid "warpqueue:;lanewalk:lw_slots=32|48|64,lw_warps=32|24;lanewalk:lw_slots=48,lw_warps=32,lw_idle_min=1|8|16" 2>&1 | tail -n 60
