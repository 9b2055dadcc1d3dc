set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_all6.log 2>&1; echo "all tests rc=$?"; tail -n 4 gpurun_out/r2_test_all6.log
timeout 600 python tools/gpu_dev.py --tag final_check --reps 5 --configs rgbbox:1000:1000:64,irreg:1000:1000:64,rgbbox:1000:1000:1,irreg:1000:1000:1,rgbbox:200:200:1,rgbbox:500:500:64 --grid "warpqueue:;warpqueue:wq_k=1" 2>&1 | grep config
python tools/trace_tail.py rgbbox 2>&1 | grep "probes0x8" | cut -c1-260
