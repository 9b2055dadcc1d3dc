"""Pins the oracle: it must reproduce the reference's own golden images bit-for-bit and the
known-answer hashes/structure facts recorded in SURVEY.md §8c, before any GPU result is compared to it."""
import hashlib

import numpy as np
import pytest

# SHA-256 over little-endian int32[h][w]; the 500^2 rows equal the reference PNGs.
KAT_SHA = {
    ("rgbbox", 200): "bce592f31741e444ee4d34b54c0a805d1019bc1a339877da40e3b3561cb0be40",
    ("rgbbox", 500): "bde03722df811efeedca73eb981ccdd9b9faf6edf9d00dc4394d3583f511a057",
    ("irreg", 200): "3ee8464286f866bec8a8011ecc7ddfe3f2701fab684a3f2bdddab5c93f3fe294",
    ("irreg", 500): "b3728459f55ca910a5555d0d8bbbbd0ae8e3f786961a2ea527c55d71720181e9",
}
KAT_SHA_1000 = {
    "rgbbox": "723bbc1045e5ccde39a0c7e828635e3ced3ddac86abd2381da6a75a830cf9535",
    "irreg": "007736b76d3011887eb12b63b8136a827424f300f0ffb4b77f16ebf8f2b48864",
}
# work counters of the reference traversal at 1000^2 (SURVEY.md §8c / BASELINE.md §3)
KAT_WORK_1000 = {
    "rgbbox": dict(segments=4022099, iterations=256792968, box_tests=117685724, leaf_tests=25443619),
    "irreg": dict(segments=1728608, iterations=109419663, box_tests=50741777, leaf_tests=9664717),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<i4").tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["rgbbox", "irreg"])
def test_oracle_reproduces_reference_png(oracle, golden, name):
    want, meta = golden[f"{name}_500"]
    assert sha(want) == meta["sha256_le_i32"]
    got, _, _ = oracle.render_scene(name, 500, 500)
    assert int((got != want).sum()) == 0, "oracle differs from the reference's golden image"


@pytest.mark.parametrize("name,size", sorted(KAT_SHA))
def test_oracle_known_answer_hashes(oracle, name, size):
    got, _, _ = oracle.render_scene(name, size, size)
    assert sha(got) == KAT_SHA[(name, size)]


@pytest.mark.parametrize("name", ["rgbbox", "irreg"])
def test_oracle_1000_hash_and_work_counters(oracle, name):
    got, _, cnt = oracle.render_scene(name, 1000, 1000)
    assert sha(got) == KAT_SHA_1000[name]
    assert cnt == KAT_WORK_1000[name]


def test_spot_pixels(oracle):
    for size in (200, 500):
        p, _, _ = oracle.render_scene("rgbbox", size, size)
        assert p[0, 0] == 0xC70000 and p[size // 2, size // 2] == 0x92BE00
        p, _, _ = oracle.render_scene("irreg", size, size)
        assert p[0, 0] == 0xA1C7FF and p[size // 2, size // 2] == 0x90BDFF


def tree_depth(left, right):
    depth = np.zeros(len(left), np.int64)
    order = [0]
    for k in order:
        for c in (left[k], right[k]):
            if c >= 0:
                depth[c] = depth[k] + 1
                order.append(int(c))
    return int(depth.max()) + 1 if len(left) else 0  # + the leaf level


def test_structure_kats(oracle):
    # rgbbox: 29 duplicate Morton codes, 10 sweeps; irreg: 0 duplicates, 15 sweeps, root box
    pr = oracle.Scene.rgbbox().prepare(500, 500)
    d = pr.dump()
    assert pr.sweeps == 10
    assert int((d["morton"][1:] == d["morton"][:-1]).sum()) == 29
    assert d["parent"][0] == -1 and (d["parent"][1:] >= 0).all()
    np.testing.assert_array_equal(d["boxes"][0], np.float32([-33, -33, -33, 33, 27, 27]))
    pr = oracle.Scene.irreg().prepare(500, 500)
    d = pr.dump()
    assert pr.sweeps == 15
    assert int((d["morton"][1:] == d["morton"][:-1]).sum()) == 0
    np.testing.assert_array_equal(d["boxes"][0], np.float32([-303, -3, -303, 297, 3, 297]))
    assert (np.diff(d["morton"].astype(np.int64)) >= 0).all()


def test_radix_tree_is_a_binary_tree_over_all_leaves(oracle):
    rng = np.random.default_rng(7)
    for n in (2, 3, 5, 64, 1000):
        keys = np.sort(rng.integers(0, 1 << 12, n).astype(np.uint32))  # many duplicates
        left = np.empty(n - 1, np.int32); right = np.empty(n - 1, np.int32); parent = np.empty(n - 1, np.int32)
        oracle.lib().oracle_radix_tree(keys.ctypes.data, n, left.ctypes.data, right.ctypes.data, parent.ctypes.data)
        leaves = sorted([~c for c in np.concatenate([left, right]) if c < 0])
        assert leaves == list(range(n))
        inner = sorted([c for c in np.concatenate([left, right]) if c >= 0])
        assert inner == list(range(1, n - 1))
        assert parent[0] == -1
        for k in range(n - 1):
            for c in (left[k], right[k]):
                if c >= 0:
                    assert parent[c] == k


# The reference's only in-tree tests for this setup path: lib/github.com/diku-dk/sorts/radix_sort_tests.fut.
# The by-key wrapper must give the permutation of a STABLE sort (sort_perm_* at :28-46).
@pytest.mark.parametrize("keys,perm", [
    ([5, 4, 3, 2, 1], [4, 3, 2, 1, 0]),
    ([5, 4, 3, 3, 2, 1], [5, 4, 2, 3, 1, 0]),
    ([1, 1, 1, 1], [0, 1, 2, 3]),
    ([], []),
])
def test_sort_perm_is_stable(oracle, keys, perm):
    k = np.asarray(keys, np.uint32)
    out = np.empty(len(keys), np.int32)
    oracle.lib().oracle_sort_perm(k.ctypes.data, len(keys), out.ctypes.data)
    assert out.tolist() == perm


def test_morton_and_primitives(oracle):
    L = oracle.lib()
    assert L.oracle_morton_3d(0.0, 0.0, 0.0) == 0
    assert L.oracle_morton_3d(1.0, 1.0, 1.0) == 0x3FFFFFFF
    assert L.oracle_morton_3d(float("nan"), 0.0, 0.0) == 0          # 0/0 axis -> coordinate 0 (irreg's y axis)
    assert L.oracle_morton_3d(1.0 / 1024, 0.0, 0.0) == 4 and L.oracle_morton_3d(0.0, 1.0 / 1024, 0.0) == 2
    box = np.float32([-1, -1, -1, 1, 1, 1])
    hit = lambda o, d: L.oracle_aabb_hit(box.ctypes.data, np.float32(list(o) + list(d)).ctypes.data)
    assert hit((0, 0, -5), (0, 0, 1)) == 1
    assert hit((0, 0, -5), (0, 0, -1)) == 0
    assert hit((2, 0, -5), (0, 0, 1)) == 0
    assert hit((1, 0, -5), (0, 0, 1)) == 1   # origin on the x=max plane with dir.x = 0: 0*inf = NaN is ignored by fmax/fmin
    s = np.float32([0, 0, 0, 1, 0.5, 0.25, 2])
    out = np.zeros(10, np.float32)
    r = np.float32([0, 0, -10, 0, 0, 1])
    assert L.oracle_sphere_hit(s.ctypes.data, r.ctypes.data, 0.1, 1e9, out.ctypes.data) == 1
    assert out[0] == 8 and tuple(out[4:7]) == (0, 0, -1) and tuple(out[7:10]) == (1, 0.5, 0.25)
    r = np.float32([0, 0, 0, 0, 0, 1])  # from the centre: root1 < t_min -> root2
    assert L.oracle_sphere_hit(s.ctypes.data, r.ctypes.data, 0.1, 1e9, out.ctypes.data) == 1 and out[0] == 2
    assert L.oracle_sphere_hit(s.ctypes.data, r.ctypes.data, 0.1, 2.0, out.ctypes.data) == 0  # strict t < t_max


def test_spp1_is_the_reference_and_row_sampling(oracle):
    pr = oracle.Scene.rgbbox().prepare(64, 64)
    full, rgb, _ = pr.render(64, 64, want_rgb=True)
    part, _, _ = pr.render(64, 64, row_start=3, row_step=8)
    rows = np.arange(3, 64, 8)
    np.testing.assert_array_equal(part[rows], full[rows])
    assert (np.delete(part, rows, axis=0) == 0).all()
    q = (255.99 * rgb.astype(np.float32)).astype(np.float32).astype(np.int32)
    np.testing.assert_array_equal((q[..., 0] << 16) | (q[..., 1] << 8) | q[..., 2], full)
    one, _, _ = pr.render(64, 64, threads=1)
    np.testing.assert_array_equal(one, full)


def test_per_pixel_cost_sums_to_the_counters_and_shards_evenly(oracle):
    """oracle_render_cost (per-pixel bvh_fold iterations / objs_hit calls) is the same loop as oracle_render: its sums
    equal the aggregate counters, also at spp > 1; and the tile -> rank rule of the multi-GPU path (8x4 tiles,
    t % world) spreads that work evenly (the property behind 'no load balancing beyond interleaving')."""
    from raytracers_b200 import distributed as D
    for name, h, w, spp in (("rgbbox", 64, 96, 1), ("irreg", 50, 70, 3)):
        pr = getattr(oracle.Scene, name)().prepare(h, w)
        _, _, cnt = pr.render(h, w, spp=spp)
        it, seg = pr.render_cost(h, w, spp=spp)
        assert int(it.sum()) == cnt["iterations"] and int(seg.sum()) == cnt["segments"]
        assert seg.min() >= spp
    h = w = 400
    it, _ = oracle.Scene.irreg().prepare(h, w).render_cost(h, w)
    for world in (2, 4, 8):
        per_rank = [int(D.extract_rank_tiles(it, r, world).sum()) for r in range(world)]
        assert sum(per_rank) == int(it.sum())
        assert max(per_rank) / (sum(per_rank) / world) < 1.05, (world, per_rank)


def test_oracle_frame_hashes_file_is_current(oracle):
    """tests/golden/oracle_frame_hashes.json (full-frame known answers the GPU tests and bench.py compare with; written by
    tools/make_oracle_hashes.py) still is what the oracle renders - checked on its cheapest entry, irreg 4000x4000 at 1 spp
    (the north-star frame; 27.7 M segments, a few seconds); the other entries are checked for shape and format."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_frame_hashes.json")) as f:
        known = json.load(f)
    assert set(known) >= {"rgbbox_1000x1000_64spp", "irreg_1000x1000_64spp", "irreg_4000x4000_1spp", "random1M_2000x2000_2spp"}
    got, _, cnt = oracle.render_scene("irreg", 4000, 4000)
    want = known["irreg_4000x4000_1spp"]
    assert sha(got) == want["sha256_le_i32"] and cnt["segments"] == want["segments"] == 27663974
    assert all(len(v["sha256_le_i32"]) == 64 and v["shape"] == [int(x) for x in k.split("_")[1].split("x")] for k, v in known.items())
