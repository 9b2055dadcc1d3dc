#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference's own golden images.

The reference (athas/raytracers) has no tests for the ray tracer; its only fixtures are the two
README illustrations rgbbox.png / irreg.png (500x500) in the repo root, which are outputs of the
Futhark program (futhark/Makefile:29-33 targets, scaled to 500).  This script converts them to the
packed-i32 [h][w] layout futhark_entry_render returns (ray.fut:158-162: 0x00RRGGBB) and stores them
compressed, so the GPU box (which has no /root/reference) can check against them.

Run in the authoring container only:  python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

REF = os.environ.get("RAY_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def png_to_packed(path):
    img = np.asarray(Image.open(path).convert("RGB"), dtype=np.int32)  # irreg.png is palettised
    return ((img[..., 0] << 16) | (img[..., 1] << 8) | img[..., 2]).astype("<i4")


def main():
    os.makedirs(OUT, exist_ok=True)
    meta = {}
    for name in ("rgbbox", "irreg"):
        packed = png_to_packed(os.path.join(REF, f"{name}.png"))
        np.savez_compressed(os.path.join(OUT, f"{name}_500.npz"), pixels=packed)
        meta[f"{name}_500"] = {
            "shape": list(packed.shape),
            "sha256_le_i32": hashlib.sha256(packed.tobytes()).hexdigest(),
            "source": f"{name}.png (reference repo root; README.md:21,25)",
        }
    with open(os.path.join(OUT, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    sys.exit(main())
