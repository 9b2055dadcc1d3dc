# Scaling run on one multi-GPU box: NGPU=2|4|8 gpurun --gpus $NGPU -- bash tools/gpu_scale.sh  (results -> gpurun_out/)
set -x
N=${NGPU:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_final_n${N}.json 2> gpurun_out/r2_scale_final_n${N}.err; echo "bench rc=$?"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2_scale_final_n${N}.err | tail -n 5
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_scale_final_n${N}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','strict_steps','e2e','gpu_launches','shard_kernel_ms_per_rank')})
PY
if [ "$N" = "8" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --workload irreg4000 > gpurun_out/r2_scale_final_irreg4000_n${N}.json 2> gpurun_out/r2_scale_final_irreg4000_n${N}.err; echo "irreg4000 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_scale_final_irreg4000_n${N}.json').read().strip().splitlines()[-1])
print('irreg4000', {k:d[k] for k in ('value','ms_per_step','strict_steps','e2e')})
PY
fi
