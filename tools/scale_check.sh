# usage: tools/scale_check.sh N   (on a box with N GPUs) -> gpurun_out/scale_nN*.json
N=${1:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@"; }
run > gpurun_out/scale_n${N}.json 2> gpurun_out/scale_n${N}.err || tail -c 1500 gpurun_out/scale_n${N}.err
run --workload irreg4000 --steps 5 > gpurun_out/scale_irreg4000_n${N}.json 2> gpurun_out/scale_irreg4000_n${N}.err || tail -c 1500 gpurun_out/scale_irreg4000_n${N}.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_*n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"], d["shard_kernel_ms_per_rank"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
