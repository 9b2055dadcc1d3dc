import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import raytracers_b200 as rb
for scene_name in ("rgbbox", "irreg"):
    for world in (1, 8):
        for w in (24, 28, 32):
            with rb.Context(wq_warps=w) as ctx:
                ctx.trace_warps(True)
                sc = ctx.scene(scene_name); prep = ctx.prepare_scene(1000, 1000, sc)
                ctx.set_shard(0, world)
                tiles = torch.zeros((251 * 126 * 32 // world + 64,), dtype=torch.int32, device="cuda")
                t_end = time.time() + 0.3
                while time.time() < t_end:
                    ctx.render_shard_into(tiles.data_ptr(), 1000, 1000, prep, spp=64); torch.cuda.synchronize()
                rec = []
                for _ in range(5):
                    ctx.render_shard_into(tiles.data_ptr(), 1000, 1000, prep, spp=64); torch.cuda.synchronize()
                    rec.append((ctx.last_render_ms(), np.sort(ctx.warp_trace())))
                ms, t = sorted(rec, key=lambda r: r[0])[2]
                print(scene_name, "world", world, "warps", w, "ms", round(ms, 3), "exit p1/p50/p99/max", [round(float(np.percentile(t, p)) / 1e3, 3) for p in (1, 50, 99, 100)], flush=True)
