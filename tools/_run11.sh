set -x
timeout 600 python tools/gpu_dev.py --tag k2_sweep --reps 5 --configs rgbbox:1000:1000:64,irreg:1000:1000:64 \
  --grid "warpqueue:;warpqueue:wq_k=2,wq_warps=20|24,wq_ncap=256|512" 2>&1 | grep config
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_all5.log 2>&1; echo "all tests rc=$?"; tail -n 4 gpurun_out/r2_test_all5.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2_bench3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','strict_steps','e2e','frame_ms_1spp','gpu_launches')})
print(d['roofline']['per_scene'], d['roofline']['frac'], d['roofline_issue']['frac'], d['roofline_issue']['per_scene'])
print({k:(v['ms_per_frame'],v['roofline_frac'],v['issue_frac'],v['parity']['differing'] if v['parity'] else None) for k,v in d['extra'].items()})
print(d['cpu_baseline']['value'], d['parity']['differing'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench3_ref.json 2>&1; tail -c 600 gpurun_out/r2_bench3_ref.json
