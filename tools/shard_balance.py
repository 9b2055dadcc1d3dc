"""How evenly does a tile -> GPU assignment spread a frame?  CPU only: per-pixel work (bvh_fold iterations) from the
oracle, summed per rank for the assignment rules under study.  Prints max/mean work per rank (1.0 = perfect)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as O  # noqa: E402

TW, TH = 8, 4


def tile_costs(name, h, w):
    it, _ = getattr(O.Scene, name)().prepare(h, w).render_cost(h, w)
    ty, tx = (h + TH - 1) // TH, (w + TW - 1) // TW
    pad = np.zeros((ty * TH, tx * TW), np.int64)
    pad[:h, :w] = it
    return pad.reshape(ty, TH, tx, TW).sum(axis=(1, 3)).reshape(-1)  # row-major tile order


def mix(x):  # the 32-bit finaliser the kernels would use
    x = np.asarray(x, np.uint64)
    x = (x ^ (x >> np.uint64(16))) * np.uint64(0x7feb352d) & np.uint64(0xffffffff)
    x = (x ^ (x >> np.uint64(15))) * np.uint64(0x846ca68b) & np.uint64(0xffffffff)
    return x ^ (x >> np.uint64(16))


def rules(n, world):
    t = np.arange(n)
    blk, r = t // world, t % world
    return {"t % world": r, "rotate by hash(block)": (r + mix(blk) % np.uint64(world)).astype(np.int64) % world,
            "rotate by block": (r + blk) % world}


def main():
    h = w = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    for name in ("rgbbox", "irreg"):
        c = tile_costs(name, h, w)
        for world in (2, 4, 8):
            for rule, owner in rules(c.size, world).items():
                per = np.bincount(owner, weights=c, minlength=world)
                print(f"{name} {h}x{w} world={world} {rule:24s} max/mean = {per.max() / per.mean():.4f}  min/mean = {per.min() / per.mean():.4f}")


if __name__ == "__main__":
    main()
