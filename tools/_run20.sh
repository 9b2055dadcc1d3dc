set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_all9.log 2>&1; echo "all tests rc=$?"; tail -n 4 gpurun_out/r2_test_all9.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 7
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2_bench5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench5.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','strict_steps','e2e','frame_ms_1spp','gpu_launches')})
print(d['roofline']['per_scene'], d['roofline']['frac'], d['roofline_issue']['frac'], d['roofline_issue']['per_scene'])
print({k:(v['ms_per_frame'],v['roofline_frac'],v['issue_frac'],v['parity']['differing'] if v['parity'] else None) for k,v in d['extra'].items()})
print(d['cpu_baseline']['value'], d['parity']['differing'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench5_ref.json 2>&1; tail -c 300 gpurun_out/r2_bench5_ref.json
for s in rgbbox irreg; do ./examples/_built/main_ref -s $s -n 1000 -m 1000 2>&1 | grep -i "in 0"; done
