set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "peer_frame or context or batch" 2>&1 | tail -n 12
timeout 900 python tools/gpu_dev.py --tag stagecap --reps 5 --configs random:2000:2000:2:1000000 \
  --grid "warpqueue:stage_cap=2048|16384|65536|100000000,wq_warps=16|24|32" 2>&1 | grep config
timeout 600 python tools/gpu_dev.py --tag stagecap_irreg --reps 5 --configs irreg:1000:1000:64,irreg:4000:4000:1 \
  --grid "warpqueue:stage_cap=100000000,wq_warps=32;warpqueue:stage_cap=8192|32768,wq_warps=24|28" 2>&1 | grep config
