#!/usr/bin/env python
"""Minimal launch sequence for ncu: N frames of one scene with one kernel variant.  Never a bench number."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracers_b200 as R  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="rgbbox")
ap.add_argument("--size", type=int, default=1000)
ap.add_argument("--spp", type=int, default=1)
ap.add_argument("--kernel", default="persistent")
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--tuning", default="")
a = ap.parse_args()
tuning = dict((k, int(v)) for k, v in (kv.split("=") for kv in a.tuning.split(",") if kv))
with R.Context(kernel=a.kernel, **tuning) as ctx:
    pr = ctx.prepare_scene(a.size, a.size, ctx.scene(a.scene, n=a.n))
    import json
    print("WORK", json.dumps(dict(scene=a.scene, size=a.size, spp=a.spp, **ctx.count_work(a.size, a.size, pr, spp=a.spp))), flush=True)
    for _ in range(a.frames):
        img = ctx.render(a.size, a.size, pr, spp=a.spp)
        ctx.sync()
        print("frame ms", ctx.last_render_ms())
        img.free()
