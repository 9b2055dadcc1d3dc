set -x
timeout 900 python tools/gpu_dev.py --tag learn_order --reps 7 --configs rgbbox:1000:1000:1,irreg:1000:1000:1,irreg:4000:4000:1,rgbbox:1000:1000:64,irreg:1000:1000:64 \
  --grid "warpqueue:learn_order=0;warpqueue:learn_order=1,long_path=4|8|12|20|30" 2>&1 | grep config
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_test_all7.log 2>&1; echo "all tests rc=$?"; tail -n 4 gpurun_out/r2_test_all7.log
python - <<'PY'
# futhark_context_new time: round-1 library vs now (lazy shared-memory opt-in, fewer kernels in the product build)
import subprocess, sys, os
code = "import time,sys; sys.path.insert(0,'.'); import raytracers_b200 as R; R.load_library(); c=R.Context(); c.close(); t=time.perf_counter(); c=R.Context(); dt=time.perf_counter()-t; c.close(); print('context_new_ms', round(dt*1e3,2))"
for lib in ("raytracers_b200/_ab/libray_r1.so", ""):
    env = dict(os.environ)
    if lib: env["RAY_B200_LIB"] = os.path.abspath(lib)
    print(lib or "current", subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env).stdout.strip())
PY
