"""Host logic of the product library (scene generators, camera, LBVH build, sample offsets, shard
layout) against the oracle — no GPU needed: these entry points are pure host code."""
import numpy as np
import pytest


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name,kw", [("rgbbox", {}), ("irreg", {}), ("random", {"n": 5000, "seed": 1}), ("random", {"n": 2, "seed": 9})])
def test_scene_generators_match_reference_formulas(R, oracle, name, kw):
    s, c = R.host_scene(name, **kw)
    so, co = oracle.Scene.named(name, **kw).arrays()
    np.testing.assert_array_equal(bits(s), bits(so))
    np.testing.assert_array_equal(bits(c), bits(co))


@pytest.mark.parametrize("name,kw", [("rgbbox", {}), ("irreg", {}), ("random", {"n": 20000, "seed": 3})])
def test_lbvh_is_bit_identical_to_the_oracle(R, oracle, name, kw):
    s, c = R.host_scene(name, **kw)
    got = R.host_lbvh(s)
    pr = oracle.Scene.custom(s, c).prepare(300, 500)
    want = pr.dump()
    for k in ("morton", "perm", "left", "right", "parent"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    np.testing.assert_array_equal(bits(got["boxes"]), bits(want["boxes"]))
    assert got["refit_sweeps"] == pr.sweeps
    np.testing.assert_array_equal(bits(R.host_camera(c, 300, 500)), bits(want["cam"]))


def test_structure_facts(R):
    s, _ = R.host_scene("rgbbox")
    t = R.host_lbvh(s)
    # 10 sweeps < tree height => 3 stale inner boxes, kept on purpose
    assert (t["refit_sweeps"], t["stale_nodes"]) == (10, 3)
    s, _ = R.host_scene("irreg")
    t = R.host_lbvh(s)
    assert (t["refit_sweeps"], t["stale_nodes"]) == (15, 0)


def test_degenerate_inputs(R, oracle):
    with pytest.raises(R.RayError):
        R.host_lbvh(np.zeros((1, 7), np.float32))          # n < 2 (bvh.fut:65)
    with pytest.raises(R.RayError):
        R.host_lbvh(np.zeros((0, 7), np.float32))
    # all spheres coincident: every Morton code equal (0/0 -> NaN -> 0), split purely by index
    s = np.tile(np.float32([1, 2, 3, 1, 1, 1, 0.5]), (37, 1))
    got = R.host_lbvh(s)
    want = oracle.Scene.custom(s, np.float32([0, 0, 10, 0, 0, 0, 60])).prepare(10, 10).dump()
    assert (got["morton"] == 0).all() and (got["perm"] == np.arange(37)).all()
    for k in ("left", "right", "parent"):
        np.testing.assert_array_equal(got[k], want[k])
    np.testing.assert_array_equal(bits(got["boxes"]), bits(want["boxes"]))
    # exactly two spheres
    s2 = np.float32([[0, 0, 0, 1, 0, 0, 1], [5, 0, 0, 0, 1, 0, 2]])
    g2 = R.host_lbvh(s2)
    assert g2["left"].tolist() == [~0] and g2["right"].tolist() == [~1] and g2["parent"].tolist() == [-1]


def test_sample_offsets(R, oracle):
    import ctypes as C
    t = R.host_sample_offsets(64)
    assert t[0, 0] == 0 and t[0, 1] == 0 and (t >= 0).all() and (t < 1).all()
    for s in (1, 7, 63):
        ox, oy = C.c_float(), C.c_float()
        oracle.lib().oracle_sample_offset(s, C.byref(ox), C.byref(oy))
        assert (t[s, 0], t[s, 1]) == (ox.value, oy.value)


@pytest.mark.parametrize("h,w,world", [(8, 8, 1), (200, 200, 2), (37, 53, 4), (1000, 1000, 8), (4, 8, 8), (1, 1, 3)])
def test_shard_layout_matches_the_library(R, h, w, world):
    from raytracers_b200 import distributed as D
    L = R.load_library()
    _, _, n_tiles, padded = D.tile_layout(h, w, world)
    assert L.ray_b200_shard_tiles_padded(h, w, world) == padded
    counts = [L.ray_b200_shard_tiles(h, w, r, world) for r in range(world)]
    assert counts == [D.rank_tile_count(h, w, r, world) for r in range(world)]
    assert sum(counts) == n_tiles and max(counts) <= padded
    # tiling round trip on the CPU statements of render_shard / detile
    img = np.arange(h * w, dtype=np.int32).reshape(h, w) + 1
    gathered = np.stack([D.extract_rank_tiles(img, r, world) for r in range(world)])
    np.testing.assert_array_equal(D.detile_reference(gathered, h, w, world), img)


def _adversarial_scene(rng, n, mode):
    s = np.zeros((n, 7), np.float32)
    if mode == "grid":        # few distinct centres -> many duplicate Morton codes (tie-break 32 + clz(i ^ j))
        s[:, :3] = rng.integers(0, 3, size=(n, 3)).astype(np.float32)
    elif mode == "line":      # two degenerate axes: (max - min) = 0 -> 0/0 = NaN -> Morton coordinate 0
        s[:, 0] = rng.random(n, dtype=np.float32) * 100
    elif mode == "plane":
        s[:, [0, 2]] = rng.random((n, 2), dtype=np.float32) * 50 - 25
    elif mode == "huge":      # extreme magnitudes: clamp(x * 1024, 0, 1023) and float rounding in the normalisation
        s[:, :3] = (rng.random((n, 3), dtype=np.float32) - 0.5) * np.float32(1e30)
    else:
        s[:, :3] = rng.standard_normal((n, 3)).astype(np.float32) * 10
    s[:, 3:6] = rng.random((n, 3), dtype=np.float32)
    s[:, 6] = rng.random(n, dtype=np.float32) * 2 + np.float32(0.01)
    return s


@pytest.mark.parametrize("mode", ["grid", "line", "plane", "huge", "cloud"])
def test_lbvh_matches_oracle_on_adversarial_scenes(R, oracle, mode):
    """Host builder == oracle, array for array and bit for bit, on inputs chosen to stress S14-S18: duplicate keys,
    NaN Morton axes, clamping, stale refits at awkward n (powers of two +-1)."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(mode.encode()))
    cam = np.float32([0, 0, 50, 0, 0, 0, 60])
    for n in (2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 100, 257, 1000):
        s = _adversarial_scene(rng, n, mode)
        got = R.host_lbvh(s)
        want_pr = oracle.Scene.custom(s, cam).prepare(16, 16)
        want = want_pr.dump()
        for k in ("morton", "perm", "left", "right", "parent"):
            np.testing.assert_array_equal(got[k], want[k], err_msg=f"{mode} n={n} {k}")
        np.testing.assert_array_equal(bits(got["boxes"]), bits(want["boxes"]), err_msg=f"{mode} n={n} boxes")
        assert got["refit_sweeps"] == want_pr.sweeps
        # structural invariants of a Karras tree: every leaf and every inner node except the root has exactly one parent
        refs = np.concatenate([got["left"], got["right"]])
        leaves = np.sort(~refs[refs < 0])
        inner = np.sort(refs[refs >= 0])
        np.testing.assert_array_equal(leaves, np.arange(n))
        np.testing.assert_array_equal(inner, np.arange(1, n - 1))
