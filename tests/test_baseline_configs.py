"""Oracle parity AT the BASELINE.json configs themselves, with the DEFAULT kernel (`RAY_KERNEL=auto`).

The small-frame tests of test_gpu_parity.py compare every kernel with the oracle; the sizes BASELINE.json quotes its
metric on are too slow for a full CPU render inside a test, so here the oracle renders a row sample of exactly those
frames (`oracle_render(row_start, row_step)`: rows j with (j - row_start) % row_step == 0; every sample of those pixels)
and the same rows of the GPU frame must be bit-identical.  Reference semantics: ray.fut:150-169 (trace_ray / render_image),
ray.fut:246-247 (entry render); the spp rule is the extension of SURVEY.md §8d (offset (0,0) at sample 0)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rows_equal(got, want, row_start, row_step, what):
    rows = np.arange(row_start, got.shape[0], row_step)
    g, w = got[rows], want[rows]
    bad = int((g != w).sum())
    if bad:
        d = np.abs(((g[..., None] >> np.array([16, 8, 0])) & 255) - ((w[..., None] >> np.array([16, 8, 0])) & 255)).max()
        raise AssertionError(f"{what}: {bad} of {g.size} sampled pixels differ (max channel delta {d})")
    return rows.size


def _gpu(R, scene_args, h, w, spp, **ctx_kw):
    with R.Context(**ctx_kw) as ctx:
        sc = ctx.scene(scene_args[0], n=scene_args[1] if len(scene_args) > 1 else None,
                       seed=scene_args[2] if len(scene_args) > 2 else 1)
        pr = ctx.prepare_scene(h, w, sc)
        out = ctx.render_host(h, w, pr, spp=spp)
        pr.free(); sc.free()
    return out


@pytest.mark.parametrize("name", ["rgbbox", "irreg"])
def test_headline_1000x1000_64spp_vs_oracle_rows(R, oracle, name):
    """BASELINE configs[1] / [2]: 1000x1000 at 64 spp, default kernel, rows j % 8 == 0 (125 rows, 8 M samples)."""
    h = w = 1000
    want, _, _ = oracle.Scene.named(name).prepare(h, w).render(h, w, spp=64, row_start=0, row_step=8)
    got = _gpu(R, (name,), h, w, 64)
    assert _rows_equal(got, want, 0, 8, f"{name} 1000^2 64 spp auto") == 125


def test_irreg_4000x4000_1spp_full_frame_vs_oracle(R, oracle):
    """The north-star target config at 1 spp: the WHOLE 4000x4000 frame (27.7 M segments) against the oracle, default
    kernel (this is the size where `auto` switches the packet walk on) and the plain warp-queue kernel."""
    h = w = 4000
    want, _, cnt = oracle.Scene.irreg().prepare(h, w).render(h, w)
    assert cnt["segments"] == 27663974
    for kernel in ("auto", "warpqueue"):
        got = _gpu(R, ("irreg",), h, w, 1, kernel=kernel)
        _rows_equal(got, want, 0, 1, f"irreg 4000^2 1 spp {kernel}")


def test_irreg_4000x4000_256spp_vs_oracle_rows(R, oracle):
    """BASELINE configs[3]: irreg 4000x4000 at 256 spp, default kernel; 16 rows (j % 256 == 128: sky rows and the
    50-bounce ground rows alike), 16 M samples on the oracle."""
    h = w = 4000
    want, _, _ = oracle.Scene.irreg().prepare(h, w).render(h, w, spp=256, row_start=128, row_step=256)
    got = _gpu(R, ("irreg",), h, w, 256)
    assert _rows_equal(got, want, 128, 256, "irreg 4000^2 256 spp auto") == 16


def test_million_spheres_2000x2000_16spp_vs_oracle_rows(R, oracle):
    """BASELINE configs[4]: 1 M random spheres (SURVEY §8d generator, seed 1), 2000x2000 at 16 spp, default kernel,
    device LBVH build; the oracle builds the same tree on the CPU and renders 8 rows (j % 250 == 125)."""
    h = w = 2000
    n = 1000000
    pr = oracle.Scene.random(n, 1).prepare(h, w)
    want, _, _ = pr.render(h, w, spp=16, row_start=125, row_step=250)
    got = _gpu(R, ("random", n, 1), h, w, 16)
    assert _rows_equal(got, want, 125, 250, "1M spheres 2000^2 16 spp auto") == 8
    assert len(np.unique(got)) > 1000


def test_large_frame_all_kernels_agree_with_default(R):
    """4000x4000 irreg: the default kernel and the explicit warp-queue kernel against the lane-bound anchor kernels."""
    import hashlib

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a, dtype="<i4").tobytes()).hexdigest()
    ref = sha(_gpu(R, ("irreg",), 4000, 4000, 1, kernel="mega"))
    for kernel in ("auto", "warpqueue"):
        assert sha(_gpu(R, ("irreg",), 4000, 4000, 1, kernel=kernel)) == ref, kernel
