"""EVERY pixel of whole frames at benchmark size against the oracle's known answers (default kernel).

The oracle needs minutes per frame at these sizes (11 minutes for BASELINE.json configs[3] and [4] each), so it is not run
here: tools/make_oracle_hashes.py rendered the frames once in the authoring container and recorded the SHA-256 of the
little-endian int32[h][w] arrays in tests/golden/oracle_frame_hashes.json; tests/test_oracle_golden.py re-renders the cheapest
entry on the CPU to keep the file honest.  Row samples of the same configs against a LIVE oracle are in
test_baseline_configs.py.  Reference semantics: ray.fut:150-169, 246-247; spp > 1 is the extension of SURVEY.md section 8d."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu(R, scene_args, h, w, spp, **ctx_kw):
    with R.Context(**ctx_kw) as ctx:
        sc = ctx.scene(scene_args[0], n=scene_args[1] if len(scene_args) > 1 else None,
                       seed=scene_args[2] if len(scene_args) > 2 else 1)
        pr = ctx.prepare_scene(h, w, sc)
        out = ctx.render_host(h, w, pr, spp=spp)
        pr.free(); sc.free()
    return out


@pytest.mark.parametrize("key,scene,h,w,spp", [("rgbbox_1000x1000_64spp", ("rgbbox",), 1000, 1000, 64),
                                               ("irreg_1000x1000_64spp", ("irreg",), 1000, 1000, 64),
                                               ("irreg_4000x4000_1spp", ("irreg",), 4000, 4000, 1),
                                               ("rgbbox_2000x2000_16spp", ("rgbbox",), 2000, 2000, 16),
                                               ("random1M_2000x2000_2spp", ("random", 1000000, 1), 2000, 2000, 2),
                                               ("irreg_4000x4000_256spp", ("irreg",), 4000, 4000, 256),
                                               ("random1M_2000x2000_16spp", ("random", 1000000, 1), 2000, 2000, 16)])
def test_full_frame_known_answers(R, key, scene, h, w, spp):
    """EVERY pixel of whole frames at benchmark size, default kernel: the SHA-256 of the frame equals the oracle's
    (tests/golden/oracle_frame_hashes.json, written by tools/make_oracle_hashes.py - minutes of CPU, too slow to render
    here - the last two took 11 minutes each): BASELINE configs[1] / [2] (1000x1000, 64 spp), configs[3] (irreg 4000x4000
    at 256 spp, 7.08 G segments) and its 1-spp frame, configs[4] (1 M spheres, 2000x2000 at 16 spp) and its 2-spp frame,
    rgbbox 2000x2000 at 16 spp."""
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_frame_hashes.json")) as f:
        want = json.load(f)[key]
    got = _gpu(R, scene, h, w, spp)
    assert list(got.shape) == want["shape"]
    assert hashlib.sha256(np.ascontiguousarray(got, "<i4").tobytes()).hexdigest() == want["sha256_le_i32"], key
