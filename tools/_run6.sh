set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_all3.log 2>&1; echo "all tests rc=$?"; tail -n 4 gpurun_out/r2_test_all3.log
timeout 600 python tools/gpu_dev.py --tag stagecap2 --reps 5 --configs random:2000:2000:2:1000000 \
  --grid "warpqueue:;warpqueue:stage_cap=2048,wq_warps=26|28|30;warpqueue:stage_cap=1024|3072,wq_warps=32;warpqueue:wq_packet=0|16|32" 2>&1 | grep config
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','frame_ms_1spp','parity')})
print({k:(v['ms_per_frame'],v['roofline_frac'],v['issue_frac'],v['parity']['differing'] if v['parity'] else None) for k,v in d['extra'].items()})
print(d['roofline']['per_scene'], d['roofline_issue'])
PY
