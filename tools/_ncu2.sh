set -x
for cfg in "rgbbox 1000 64" "irreg 1000 64"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_warpqueue -s 1 -c 1 -f -o gpurun_out/r2b_wq_$1 python tools/profile_target.py --scene $1 --size $2 --spp $3 --kernel warpqueue --frames 2 > gpurun_out/r2b_ncu_$1.log 2>&1
  tail -n 3 gpurun_out/r2b_ncu_$1.log
done
ls -la gpurun_out | grep r2b
