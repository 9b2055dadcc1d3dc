# ncu evidence for profiles/ (round 2): launch list of the bench command + one full capture of the default render kernel per
# BASELINE scene (rgbbox / irreg at 64 spp, the 1 M-sphere scene at 2 spp), exported as details text, raw csv and the
# per-source-line csv that tools/ncu_source_summary.py condenses.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launch_list.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_under_ncu.log 2>&1
for cfg in "rgbbox 1000 64 x" "irreg 1000 64 x" "random 2000 2 1000000"; do
  set -- $cfg
  extra=""; [ "$4" != "x" ] && extra="--n $4"
  ncu --set full --clock-control none --import-source on -k regex:render_warpqueue -s 1 -c 1 -f -o gpurun_out/r2_wq_$1 python tools/profile_target.py --scene $1 --size $2 --spp $3 --kernel warpqueue --frames 2 $extra > gpurun_out/r2_ncu_$1.log 2>&1
  ncu -i gpurun_out/r2_wq_$1.ncu-rep --page details > gpurun_out/r2_wq_$1_details.txt 2>&1
  ncu -i gpurun_out/r2_wq_$1.ncu-rep --page raw --csv > gpurun_out/r2_wq_$1_raw.csv 2>&1
done
ls -la gpurun_out | tail -n 12
