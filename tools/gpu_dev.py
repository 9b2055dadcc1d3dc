#!/usr/bin/env python
"""Development harness for the GPU box: times kernel variants / tuning parameters on the BASELINE
configs and checks each frame against a hash of the first variant (and the oracle where cheap).
Writes gpurun_out/dev_<tag>.json.  Not part of the product or of bench.py."""
import argparse
import hashlib
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracers_b200 as R  # noqa: E402


WARM_S = 0.3


def run(name, h, w, spp, kernel, reps, n=None, **tuning):
    with R.Context(kernel=kernel, **tuning) as ctx:
        pr = ctx.prepare_scene(h, w, ctx.scene(name, n=n))
        ms = []
        img = None
        t_end = time.time() + WARM_S      # let the SM clocks ramp up from idle before timing
        while time.time() < t_end:
            ctx.render(h, w, pr, spp=spp).free()
            ctx.sync()
        for _ in range(reps + 2):
            if img is not None:
                img.free()
            img = ctx.render(h, w, pr, spp=spp)
            ctx.sync()
            ms.append(ctx.last_render_ms())
        pix = img.values()
    ms = sorted(ms[2:])
    return ms[len(ms) // 2], ms[0], hashlib.sha256(pix.tobytes()).hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="sweep")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--configs", default="rgbbox:1000:1000:1,irreg:1000:1000:1")
    ap.add_argument("--kernels", default="mega,persistent")
    ap.add_argument("--blocks_per_sm", default="2")
    ap.add_argument("--refill_min", default="8")
    ap.add_argument("--smem_budget", default="65536")
    ap.add_argument("--tail_from", default="8")
    ap.add_argument("--wq_warps", default="16")
    ap.add_argument("--wq_k", default="1")
    ap.add_argument("--permute", default="1")
    ap.add_argument("--wq_packet", default="0")
    ap.add_argument("--wq_refill", default="1")
    ap.add_argument("--wq_ncap", default="0")
    ap.add_argument("--grid", default="", help="generic sweep: 'kernel:key=v1|v2,key2=v3|v4;kernel2:...' (cartesian product per kernel); "
                                               "replaces --kernels and the per-kernel option lists")
    a = ap.parse_args()
    rows = []
    if a.grid:
        for cfg in a.configs.split(","):
            parts = cfg.split(":")
            name, h, w, spp = parts[0], int(parts[1]), int(parts[2]), int(parts[3])
            n = int(parts[4]) if len(parts) > 4 else None
            ref_hash = None
            for spec in a.grid.split(";"):
                kernel, _, kv = spec.partition(":")
                keys, vals = [], []
                for item in (kv.split(",") if kv else []):
                    k, _, v = item.partition("=")
                    keys.append(k)
                    vals.append([int(x) for x in v.split("|")])
                for combo in itertools.product(*vals):
                    tuning = dict(zip(keys, combo))
                    t0 = time.time()
                    try:
                        med, best, hsh = run(name, h, w, spp, kernel, a.reps, n=n, **tuning)
                    except R.RayError as e:
                        print(json.dumps(dict(config=cfg, kernel=kernel, **tuning, error=str(e))), flush=True)
                        continue
                    ref_hash = ref_hash or hsh
                    row = dict(config=cfg, kernel=kernel, **tuning, ms_median=round(med, 4), ms_best=round(best, 4), hash=hsh,
                               same_as_first=(hsh == ref_hash), wall=round(time.time() - t0, 2))
                    rows.append(row)
                    print(json.dumps(row), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"dev_{a.tag}.json"), "w") as f:
            json.dump(rows, f, indent=1)
        return
    for cfg in a.configs.split(","):
        parts = cfg.split(":")
        name, h, w, spp = parts[0], int(parts[1]), int(parts[2]), int(parts[3])
        n = int(parts[4]) if len(parts) > 4 else None
        ref_hash = None
        for kernel in a.kernels.split(","):
            if kernel == "mega":
                grid = [("x", "x", "x")]
            elif kernel == "wavefront":
                grid = itertools.product(a.blocks_per_sm.split(","), a.tail_from.split(","), a.smem_budget.split(","))
            elif kernel in ("warpqueue", "streamqueue"):
                grid = itertools.product(a.wq_warps.split(","), a.wq_ncap.split(","), a.wq_packet.split(","))
            else:
                grid = itertools.product(a.blocks_per_sm.split(","), a.refill_min.split(","), a.smem_budget.split(","))
            for bps, rf, sb in grid:
                if kernel == "mega":
                    tuning = {}
                elif kernel == "wavefront":
                    tuning = dict(blocks_per_sm=int(bps), tail_from=int(rf), smem_budget=int(sb))
                elif kernel in ("warpqueue", "streamqueue"):
                    tuning = dict(wq_warps=int(bps), wq_ncap=int(rf), wq_packet=int(sb), wq_refill=int(a.wq_refill.split(",")[0]))
                else:
                    tuning = dict(blocks_per_sm=int(bps), refill_min=int(rf), smem_budget=int(sb))
                t0 = time.time()
                try:
                    med, best, hsh = run(name, h, w, spp, kernel, a.reps, n=n, **tuning)
                except R.RayError as e:
                    print(json.dumps(dict(config=cfg, kernel=kernel, **tuning, error=str(e))), flush=True)
                    continue
                ref_hash = ref_hash or hsh
                row = dict(config=cfg, kernel=kernel, **tuning, ms_median=round(med, 4), ms_best=round(best, 4), hash=hsh,
                           same_as_first=(hsh == ref_hash), wall=round(time.time() - t0, 2))
                rows.append(row)
                print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dev_{a.tag}.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
