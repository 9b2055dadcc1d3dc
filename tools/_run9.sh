set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_all4.log 2>&1; echo "all tests rc=$?"; tail -n 6 gpurun_out/r2_test_all4.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"; tail -n 5 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','strict_steps','e2e','frame_ms_1spp','parity','gpu_launches')})
print({k:(v['ms_per_frame'],v['roofline_frac'],v['issue_frac'],v['parity']['differing'] if v['parity'] else None) for k,v in d['extra'].items()})
PY
# N4 experiment: ray re-sort between bounces (wavefront kernel), 1 M spheres and irreg 4000^2, 1 spp
timeout 900 python tools/gpu_dev.py --tag n4_resort --reps 3 --configs random:2000:2000:1:1000000,irreg:4000:4000:1 \
  --grid "warpqueue:;wavefront:wf_sort=0|1|2|4|8" 2>&1 | grep config
for srt in 0 4; do
  ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:wavefront_bounce -s 10 -c 6 --csv --log-file gpurun_out/r2_n4_resort_sort${srt}.csv \
    python tools/profile_target.py --scene random --n 1000000 --size 2000 --spp 1 --kernel wavefront --frames 2 --tuning wf_sort=${srt} > gpurun_out/r2_n4_ncu_${srt}.log 2>&1
done
ls -la gpurun_out | tail -n 5
