set -x
python -m pytest tests/test_baseline_configs.py -x -q -m gpu > gpurun_out/r2_test_baseline.log 2>&1; echo "baseline tests rc=$?" 
python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_all.log 2>&1; echo "all tests rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err; echo "bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:render_warpqueue -s 1 -c 1 -f -o gpurun_out/r2_wq_random1M python tools/profile_target.py --scene random --n 1000000 --size 2000 --spp 2 --kernel warpqueue --frames 2 > gpurun_out/r2_ncu_random1M.log 2>&1
tail -3 gpurun_out/r2_test_baseline.log gpurun_out/r2_test_all.log
cat gpurun_out/r2_bench0.json | head -c 6000
