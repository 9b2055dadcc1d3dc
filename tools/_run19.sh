set -x
python tools/context_new_time.py
for s in rgbbox irreg; do ./examples/_built/main_ref -s $s -n 1000 -m 1000 2>&1 | head -n 2; RAY_LEARN_ORDER=0 ./examples/_built/main_ref -s $s -n 1000 -m 1000 2>&1 | head -n 2; done
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import raytracers_b200 as R
for name in ("rgbbox", "irreg"):
    with R.Context() as ctx:
        pr = ctx.prepare_scene(1000, 1000, ctx.scene(name))
        ms = []
        for i in range(5):
            t = time.perf_counter(); img = ctx.render(1000, 1000, pr); ctx.sync(); ms.append(round((time.perf_counter() - t) * 1e3, 3)); img.free()
        print(name, "wall ms of frames 1..5 (render + sync):", ms)
PY
timeout 900 python tools/gpu_dev.py --tag learn_order3 --reps 3 --configs irreg:4000:4000:256,irreg:4000:4000:16,rgbbox:2000:2000:16 \
  --grid "warpqueue:learn_order=0;warpqueue:learn_order=1" 2>&1 | grep config
