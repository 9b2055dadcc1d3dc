"""ctypes binding for the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (raytracers_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


class Counters(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("iterations", C.c_uint64), ("box_tests", C.c_uint64),
                ("leaf_tests", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i64, i32, u64, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_float
        L.oracle_scene_rgbbox.restype = vp
        L.oracle_scene_irreg.restype = vp
        L.oracle_scene_random.restype = vp
        L.oracle_scene_random.argtypes = [i64, u64]
        L.oracle_scene_custom.restype = vp
        L.oracle_scene_custom.argtypes = [vp, i64, vp]
        L.oracle_scene_free.argtypes = [vp]
        L.oracle_scene_num_spheres.restype = i64
        L.oracle_scene_num_spheres.argtypes = [vp]
        L.oracle_scene_get.argtypes = [vp, vp, vp]
        L.oracle_prepare_scene.restype = vp
        L.oracle_prepare_scene.argtypes = [i64, i64, vp]
        L.oracle_prepared_free.argtypes = [vp]
        L.oracle_prepared_sweeps.restype = i32
        L.oracle_prepared_sweeps.argtypes = [vp]
        L.oracle_prepared_dump.argtypes = [vp] * 8
        L.oracle_render.restype = C.c_int
        L.oracle_render.argtypes = [vp, i64, i64, i32, vp, vp, i32, i32, i32, C.POINTER(Counters)]
        L.oracle_render_cost.restype = C.c_int
        L.oracle_render_cost.argtypes = [vp, i64, i64, i32, vp, vp, i32]
        L.oracle_num_procs.restype = C.c_int
        L.oracle_morton_3d.restype = C.c_uint32
        L.oracle_morton_3d.argtypes = [f32, f32, f32]
        L.oracle_aabb_hit.restype = C.c_int
        L.oracle_aabb_hit.argtypes = [vp, vp]
        L.oracle_sphere_hit.restype = C.c_int
        L.oracle_sphere_hit.argtypes = [vp, vp, f32, f32, vp]
        L.oracle_sample_offset.argtypes = [i32, C.POINTER(f32), C.POINTER(f32)]
        L.oracle_sort_perm.argtypes = [vp, i64, vp]
        L.oracle_radix_tree.argtypes = [vp, i64, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Scene:
    """Mirror of the reference's `scene` type (ray.fut:171-174)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: scene construction failed")
        self._h = handle

    @classmethod
    def rgbbox(cls):
        return cls(lib().oracle_scene_rgbbox())

    @classmethod
    def irreg(cls):
        return cls(lib().oracle_scene_irreg())

    @classmethod
    def random(cls, n, seed=1):
        return cls(lib().oracle_scene_random(int(n), int(seed)))

    @classmethod
    def custom(cls, spheres, cam7):
        s = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 7)
        c = np.ascontiguousarray(cam7, dtype=np.float32).reshape(7)
        return cls(lib().oracle_scene_custom(_p(s), s.shape[0], _p(c)))

    @classmethod
    def named(cls, name, **kw):
        if name == "rgbbox":
            return cls.rgbbox()
        if name == "irreg":
            return cls.irreg()
        if name.startswith("random"):
            return cls.random(kw.get("n", 1000), kw.get("seed", 1))
        raise ValueError(name)

    @property
    def num_spheres(self):
        return int(lib().oracle_scene_num_spheres(self._h))

    def arrays(self):
        n = self.num_spheres
        s = np.empty((n, 7), np.float32)
        c = np.empty(7, np.float32)
        lib().oracle_scene_get(self._h, _p(s), _p(c))
        return s, c

    def prepare(self, h, w):
        return Prepared(lib().oracle_prepare_scene(int(h), int(w), self._h), self.num_spheres)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_scene_free(self._h)
            self._h = None


class Prepared:
    """Mirror of `prepared_scene` (ray.fut:239)."""

    def __init__(self, handle, n):
        if not handle:
            raise RuntimeError("oracle: prepare_scene failed (needs >= 2 spheres)")
        self._h = handle
        self.n = n

    @property
    def sweeps(self):
        return int(lib().oracle_prepared_sweeps(self._h))

    def dump(self):
        n = self.n
        out = dict(morton=np.empty(n, np.uint32), perm=np.empty(n, np.int32), left=np.empty(n - 1, np.int32),
                   right=np.empty(n - 1, np.int32), parent=np.empty(n - 1, np.int32),
                   boxes=np.empty((n - 1, 6), np.float32), cam=np.empty(12, np.float32))
        lib().oracle_prepared_dump(self._h, _p(out["morton"]), _p(out["perm"]), _p(out["left"]), _p(out["right"]),
                                   _p(out["parent"]), _p(out["boxes"]), _p(out["cam"]))
        return out

    def render(self, h, w, spp=1, want_rgb=False, row_start=0, row_step=1, threads=0, out=None):
        """Returns (pixels int32[h][w], rgb float32[h][w][3] or None, counters dict)."""
        pix = out if out is not None else np.zeros((h, w), np.int32)
        rgb = np.zeros((h, w, 3), np.float32) if want_rgb else None
        cnt = Counters()
        rc = lib().oracle_render(self._h, int(h), int(w), int(spp), _p(pix), _p(rgb), int(row_start), int(row_step),
                                 int(threads), C.byref(cnt))
        if rc != 0:
            raise RuntimeError("oracle_render failed")
        return pix, rgb, cnt.as_dict()

    def render_cost(self, h, w, spp=1, threads=0):
        """Per-pixel work of the reference traversal: (bvh_fold iterations int32[h][w], objs_hit calls int32[h][w])."""
        it = np.zeros((h, w), np.int32)
        seg = np.zeros((h, w), np.int32)
        if lib().oracle_render_cost(self._h, int(h), int(w), int(spp), _p(it), _p(seg), int(threads)) != 0:
            raise RuntimeError("oracle_render_cost failed")
        return it, seg

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_prepared_free(self._h)
            self._h = None


def num_procs():
    return int(lib().oracle_num_procs())


def render_scene(name, h, w, spp=1, **kw):
    sc = Scene.named(name, **{k: v for k, v in kw.items() if k in ("n", "seed")})
    pr = sc.prepare(h, w)
    return pr.render(h, w, spp, **{k: v for k, v in kw.items() if k not in ("n", "seed")})
