N=${1:-8}
for mode in "" "--no-batch"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 $mode > gpurun_out/scale_n${N}${mode}.json 2> gpurun_out/scale_n${N}${mode}.err || tail -c 1500 gpurun_out/scale_n${N}${mode}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["shard_kernel_ms_per_rank"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
