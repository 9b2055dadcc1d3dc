"""Where does a frame's time go at the end?  Renders with the warp-exit trace on (ray_b200_context_trace_warps) and
prints, per config, the kernel time, percentiles of the warps' exit times and the SM time lost behind the last paths
(idle = 1 - mean(exit) / max(exit)); each config with the heavy-first claim order off and on (same pixels: the hash
of the shard is compared).  GPU only; writes gpurun_out/trace_tail.json."""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracers_b200 as rb  # noqa: E402

CONFIGS = [("rgbbox", 1000, 1000, 1, None), ("irreg", 1000, 1000, 1, None), ("rgbbox", 1000, 1000, 64, None),
           ("irreg", 1000, 1000, 64, None), ("irreg", 4000, 4000, 4, None), ("random", 2000, 2000, 2, 1000000)]
VARIANTS = [(0, 8), (1, 8), (4, 8), (1, 5), (4, 5)]  # (probe pixels per tile or 0 = off, probe segments)


def main():
    out = {}
    only = sys.argv[1:] or None
    for (scene_name, h, w, spp, n) in CONFIGS:
        if only and scene_name not in only:
            continue
        for world in (1, 8):
            ref = None
            for hf, seg in VARIANTS:
                with rb.Context(heavy_first=hf, probe_segments=seg) as ctx:
                    ctx.trace_warps(True)
                    scene = ctx.scene(scene_name, n=n)
                    prep = ctx.prepare_scene(h, w, scene)
                    ctx.set_shard(0, world)
                    tiles = torch.zeros(((h // 4 + 1) * (w // 8 + 1) * 32 // world + 64,), dtype=torch.int32, device="cuda")
                    t_end = time.time() + 0.3
                    while time.time() < t_end:
                        ctx.render_shard_into(tiles.data_ptr(), h, w, prep, spp=spp)
                        torch.cuda.synchronize()
                    rec = []
                    for _ in range(3):
                        ctx.render_shard_into(tiles.data_ptr(), h, w, prep, spp=spp)
                        torch.cuda.synchronize()
                        rec.append((ctx.last_render_ms(), np.sort(ctx.warp_trace())))
                    ms, t = sorted(rec, key=lambda r: r[0])[1]
                    hsh = hashlib.sha256(tiles.cpu().numpy().tobytes()).hexdigest()[:16]
                    ref = ref or hsh
                    q = {f"p{p}": round(float(np.percentile(t, p)) / 1e3, 3) for p in (1, 50, 90, 99)}
                    q["max"] = round(float(t[-1]) / 1e3, 3)
                    key = f"{scene_name}_{h}x{w}_{spp}spp_shard1of{world}_probes{hf}x{seg}"
                    out[key] = {"kernel_ms": round(ms, 3), "best_ms": round(min(r[0] for r in rec), 3), "warp_exit_ms": q,
                                "idle_frac": round(1.0 - float(t.mean()) / float(t[-1]), 4), "same_pixels": hsh == ref}
                    print(key, out[key], flush=True)
                    prep.free()
                    scene.free()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/trace_tail.json", "w"), indent=1)


if __name__ == "__main__":
    main()
