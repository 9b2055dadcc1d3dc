#!/usr/bin/env python
"""Known answers for whole frames: SHA-256 of the FULL frames of the configurations the GPU sweeps time - 1000x1000 at 64
samples per pixel (BASELINE.json configs[1], [2]), irreg 4000x4000 at 1 / 16 / 256 spp (configs[3]), rgbbox 2000x2000 at
16 spp and the 1 M-sphere scene at 2000x2000 at 1 / 2 / 16 spp (configs[4]) - as the CPU oracle renders them (little-endian int32[h][w], the layout futhark_values_i32_2d
returns).  tools/gpu_dev.py prints the first 64 bits of the same hash for every frame it times.

The oracle reproduces the reference's golden PNGs bit for bit (tests/test_oracle_golden.py); spp > 1 is this repo's extension,
so these frames are pinned by the oracle only.  About 27 minutes on 8 cores (the last two entries take 22 of them).  Writes tests/golden/oracle_frame_hashes.json,
which the GPU tests and bench.py (N > 1: hashes of the frames the end-to-end leg delivered to host memory) compare against.

  python tools/make_oracle_hashes.py
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402


def main():
    out = {}
    for key, name, kw, h, w, spp in (("rgbbox", "rgbbox", {}, 1000, 1000, 64), ("irreg", "irreg", {}, 1000, 1000, 64),
                                     ("irreg", "irreg", {}, 4000, 4000, 1), ("rgbbox", "rgbbox", {}, 2000, 2000, 16),
                                     ("random1M", "random", dict(n=1000000, seed=1), 2000, 2000, 2),
                                     ("irreg", "irreg", {}, 4000, 4000, 16), ("random1M", "random", dict(n=1000000, seed=1), 2000, 2000, 1),
                                     # BASELINE.json configs[3] and [4] themselves: 7.08 G and 0.28 G segments, ~11 min each on 8 cores
                                     ("irreg", "irreg", {}, 4000, 4000, 256), ("random1M", "random", dict(n=1000000, seed=1), 2000, 2000, 16)):
        t0 = time.time()
        pix, _, cnt = O.Scene.named(name, **kw).prepare(h, w).render(h, w, spp=spp)
        out[f"{key}_{h}x{w}_{spp}spp"] = {"sha256_le_i32": hashlib.sha256(np.ascontiguousarray(pix, "<i4").tobytes()).hexdigest(),
                                          "shape": [h, w], "spp": spp, "segments": int(cnt["segments"]),
                                          "source": "oracle/oracle.cpp (CPU restatement of ray.fut, bit-exact vs rgbbox.png / irreg.png at 1 spp)"}
        print(key, h, w, spp, out[f"{key}_{h}x{w}_{spp}spp"]["sha256_le_i32"], f"{time.time() - t0:.1f} s", flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "oracle_frame_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
