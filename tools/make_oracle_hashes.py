#!/usr/bin/env python
"""Known answers for the headline frames: SHA-256 of the FULL 1000x1000 frames at 64 samples per pixel (BASELINE.json
configs[1], [2]) as the CPU oracle renders them (little-endian int32[h][w], the layout futhark_values_i32_2d returns).

The oracle reproduces the reference's golden PNGs bit for bit (tests/test_oracle_golden.py); spp > 1 is this repo's extension,
so these frames are pinned by the oracle only.  About a minute on 8 cores.  Writes tests/golden/oracle_frame_hashes.json,
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
    for name in ("rgbbox", "irreg"):
        for h, w, spp in ((1000, 1000, 64),):
            t0 = time.time()
            pix, _, cnt = O.Scene.named(name).prepare(h, w).render(h, w, spp=spp)
            out[f"{name}_{h}x{w}_{spp}spp"] = {"sha256_le_i32": hashlib.sha256(np.ascontiguousarray(pix, "<i4").tobytes()).hexdigest(),
                                               "shape": [h, w], "spp": spp, "segments": int(cnt["segments"]),
                                               "source": "oracle/oracle.cpp (CPU restatement of ray.fut, bit-exact vs rgbbox.png / irreg.png at 1 spp)"}
            print(name, out[f"{name}_{h}x{w}_{spp}spp"]["sha256_le_i32"], f"{time.time() - t0:.1f} s", flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "oracle_frame_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
