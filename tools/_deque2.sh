set -x
CFG="rgbbox:1000:1000:64,irreg:1000:1000:64,rgbbox:1000:1000:1,irreg:1000:1000:1,rgbbox:2000:2000:16,random:2000:2000:2:1000000,irreg:4000:4000:1"
RAY_B200_LIB=$PWD/raytracers_b200/_ab/libray_lifo.so timeout 600 python tools/gpu_dev.py --tag lifo --reps 5 --configs $CFG --grid "warpqueue:" 2>&1 | grep config | cut -c1-150
timeout 900 python tools/gpu_dev.py --tag deque2 --reps 5 --configs $CFG --grid "warpqueue:wq_low=0|-1|96" 2>&1 | grep config | cut -c1-170
for s in rgbbox irreg; do
timeout 200 ncu --metrics sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum --clock-control none -k regex:render_warpqueue -s 1 -c 1 python tools/profile_target.py --scene $s --size 1000 --spp 64 --kernel warpqueue --frames 2 2>&1 | grep -E "icc|gcc|inst_executed|duration"
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_deque.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/r2_test_deque.log
