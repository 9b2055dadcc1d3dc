# Full single-GPU validation: GPU test suite, both bench arms, 1-spp / 64-spp frame times.  gpurun -- bash tools/gpu_validate.sh
set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_final.log 2>&1; echo "all tests rc=$?"; tail -n 4 gpurun_out/r2_test_final.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2_bench6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','strict_steps','e2e','frame_ms_1spp','gpu_launches')})
print(d['roofline']['per_scene'], d['roofline']['frac'], d['roofline_issue']['frac'], d['roofline_issue']['per_scene'])
print({k:(v['ms_per_frame'],v['roofline_frac'],v['issue_frac'],v['parity']['differing'] if v['parity'] else None) for k,v in d['extra'].items()})
print(d['cpu_baseline']['value'], d['parity']['differing'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench6_ref.json 2>&1; tail -c 200 gpurun_out/r2_bench6_ref.json
timeout 600 python tools/gpu_dev.py --tag final_1spp --reps 7 --configs rgbbox:1000:1000:1,irreg:1000:1000:1,rgbbox:1000:1000:64,irreg:1000:1000:64 --grid "warpqueue:" 2>&1 | grep config
