#!/usr/bin/env python
"""Small renders that exercise every path of the default kernels (K3 warp-queue: item queues, packet walk, sample
spreading, pixel-bound samples, two frames in flight, compact shard layout, peer-frame flags; K5 lane-walk) for
compute-sanitizer (tools/sanitize.sh).  No torch: ctypes binding + numpy only; frames are checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracers_b200 as R  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def check(got, want, what):
    bad = int((got != want).sum())
    print(f"{what}: {'ok' if bad == 0 else f'{bad} pixels differ'}", flush=True)
    return bad


def main():
    bad = 0
    h, w = 40, 72
    want = {(n, s): getattr(O.Scene, n)().prepare(h, w).render(h, w, spp=s)[0] for n in ("rgbbox", "irreg") for s in (1, 3)}
    for kernel, tuning in (("warpqueue", {}), ("warpqueue", dict(wq_packet=8)), ("warpqueue", dict(wq_spread=0)), ("warpqueue", dict(wq_low=40, wq_k=1)),
                           ("warpqueue", dict(wq_low=-1)), ("lanewalk", {})):
        with R.Context(kernel=kernel, **tuning) as ctx:
            prep = {n: ctx.prepare_scene(h, w, ctx.scene(n)) for n in ("rgbbox", "irreg")}
            for (n, s), wnt in want.items():
                bad += check(ctx.render_host(h, w, prep[n], spp=s), wnt, f"{kernel} {tuning} {n} spp={s}")
            # two frames in flight + the compact shard layout + peer-frame flags, device buffers from futhark_new_i32_2d
            lib = ctx.lib
            zeros = np.zeros((h, w), np.int32)
            imgs = [lib.futhark_new_i32_2d(ctx.handle, zeros.ctypes.data, h, w) for _ in range(3)]
            ptrs = [lib.futhark_values_raw_i32_2d(ctx.handle, im) for im in imgs]
            flags, _ = ctx.ipc_alloc(1024)
            jobs = [dict(prepared=prep["irreg"], h=h, w=w, spp=3, out_dev=ptrs[0], done_flag=flags),
                    dict(prepared=prep["rgbbox"], h=h, w=w, spp=1, out_dev=ptrs[1], done_flag=flags + 128),
                    dict(prepared=prep["rgbbox"], h=h, w=w, spp=3, out_dev=ptrs[2], wait_flag=flags, wait_value=1)]
            ctx.render_batch(jobs)
            ctx.flag_wait(flags + 128, 1)
            ctx.sync()
            for im, key in zip(imgs, (("irreg", 3), ("rgbbox", 1), ("rgbbox", 3))):
                out = np.empty((h, w), np.int32)
                lib.futhark_values_i32_2d(ctx.handle, im, out.ctypes.data)
                bad += check(out, want[key], f"{kernel} {tuning} batch {key}")
                lib.futhark_free_i32_2d(ctx.handle, im)
            bad += ctx.flag_timeouts()
            ctx.ipc_free(flags)
    with R.Context(kernel="warpqueue") as ctx:   # deep tree: global-memory nodes (256-bit loads), overflow guard of the node queue
        n, hh, ww = 30000, 32, 48
        wnt, _, _ = O.render_scene("random", hh, ww, n=n, seed=3)
        pr = ctx.prepare_scene(hh, ww, ctx.scene_random(n, 3))
        bad += check(ctx.render_host(hh, ww, pr), wnt, "warpqueue deep tree")
    print("SANITIZE_TARGET", "OK" if bad == 0 else f"FAILED ({bad})", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
