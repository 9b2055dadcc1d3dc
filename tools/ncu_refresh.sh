# ncu evidence for profiles/: launch list of the bench command + one full capture of the render kernel per headline scene
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for s in rgbbox irreg; do
  ncu --set full --clock-control none --import-source on -k regex:render_warpqueue -s 1 -c 1 -f -o gpurun_out/wq32_$s python tools/profile_target.py --scene $s --spp 64 --kernel warpqueue --frames 2 > gpurun_out/ncu_$s.log 2>&1
  ncu -i gpurun_out/wq32_$s.ncu-rep --page details > gpurun_out/wq32_${s}_details.txt 2>&1
  ncu -i gpurun_out/wq32_$s.ncu-rep --page raw --csv > gpurun_out/wq32_${s}_raw.csv 2>&1
  rm -f gpurun_out/wq32_$s.ncu-rep
done
ls -la gpurun_out | tail -8
