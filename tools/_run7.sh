set -x
nvidia-smi -L
for g in peer nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --gather $g > gpurun_out/r2_n2_$g.json 2> gpurun_out/r2_n2_$g.err; echo "bench $g rc=$?"; tail -n 3 gpurun_out/r2_n2_$g.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_n2_$g.json').read().strip().splitlines()[-1])
    print('$g', {k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','shard_kernel_ms_per_rank')})
except Exception as e: print('parse failed', e)
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_gpu" 2>&1 | tail -n 5
python tools/gpu_dev.py --tag huge_default --reps 5 --configs random:2000:2000:2:1000000 --grid "warpqueue:" 2>&1 | grep config
