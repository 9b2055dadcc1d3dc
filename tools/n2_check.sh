set -x
for mode in "" "--no-batch"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 $mode > gpurun_out/n2_batch${mode}.json 2> gpurun_out/n2${mode}.err || tail -c 1500 gpurun_out/n2${mode}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n2_batch*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["shard_kernel_ms_per_rank"])
PY
