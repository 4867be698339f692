"""ctypes binding of the CPU ORACLE (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package `yahr_b200` never does.

A scene is passed as a plain dict / object with numpy arrays (see `scene_to_desc`):
  tri_p0, tri_p1, tri_p2, tri_n0, tri_n1, tri_n2 : (N,3) float32 ; tri_material : (N,) uint32
  sph_center : (M,3) float32 ; sph_radius : (M,) float32 ; sph_material : (M,) uint32
  prim_order : None or (N+M,) uint32 with (kind<<31 | index)
  materials : (K,7) float32 ; lights : (L,6) float32 ; bvh_max_depth : int ; split_mode : int
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_triangles", C.c_uint32),
        ("tri_p0", _f32p), ("tri_p1", _f32p), ("tri_p2", _f32p),
        ("tri_n0", _f32p), ("tri_n1", _f32p), ("tri_n2", _f32p),
        ("tri_material", _u32p),
        ("n_spheres", C.c_uint32),
        ("sph_center", _f32p), ("sph_radius", _f32p), ("sph_material", _u32p),
        ("prim_order", _u32p),
        ("n_materials", C.c_uint32), ("materials", _f32p),
        ("n_lights", C.c_uint32), ("lights", _f32p),
        ("bvh_max_depth", C.c_int32), ("split_mode", C.c_int32),
        ("n_area_lights", C.c_uint32), ("area_lights", _f32p),
    ]


class Camera(C.Structure):
    _fields_ = [("imW", C.c_float), ("imH", C.c_float), ("focalLength", C.c_float),
                ("lookDir", C.c_float * 3), ("upDir", C.c_float * 3), ("position", C.c_float * 3)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("n_node", "n_tri", "n_tricand", "n_sph", "n_shade", "n_light", "n_primary", "n_shadow",
                 "n_secondary")] + [("seconds", C.c_double), ("threads", C.c_int32), ("tiles", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    """Compile the oracle with its Makefile (g++ -O2 -ffp-contract=off, OpenMP)."""
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in ("yahr_oracle.hpp", "oracle_capi.cpp", "Makefile")):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.yo_scene_create.restype = C.c_void_p
        L.yo_scene_create.argtypes = [C.POINTER(SceneDesc)]
        L.yo_scene_destroy.argtypes = [C.c_void_p]
        for n in ("yo_bvh_num_nodes", "yo_bvh_num_prims"):
            getattr(L, n).restype = C.c_uint32
            getattr(L, n).argtypes = [C.c_void_p]
        L.yo_bvh_depth.restype = C.c_int32
        L.yo_bvh_depth.argtypes = [C.c_void_p]
        L.yo_bvh_order.argtypes = [C.c_void_p, _u32p]
        L.yo_bvh_dump_preorder.restype = C.c_uint32
        L.yo_bvh_dump_preorder.argtypes = [C.c_void_p, _u32p, _u32p, _u32p, _f32p]
        L.yo_num_batches.restype = C.c_int64
        L.yo_num_batches.argtypes = [C.c_int64] * 3
        L.yo_batch_window.argtypes = [C.c_int64] * 4 + [C.POINTER(C.c_int32)]
        L.yo_render.restype = C.c_int
        L.yo_render.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_int, C.c_int, C.c_uint64, _f32p, _u32p, _f32p,
                                C.POINTER(Stats), C.c_int, C.c_int, C.c_int, C.c_int64]
        L.yo_cross.argtypes = [_f32p] * 3
        L.yo_norm.argtypes = [_f32p] * 2
        L.yo_collide_all_const.restype = C.c_int
        L.yo_collide_all_const.argtypes = [C.c_int, C.POINTER(C.c_int), _f32p, C.POINTER(C.c_int)]
        L.yo_camera_ray.argtypes = [C.POINTER(Camera), C.c_float, C.c_float, _f32p]
        L.yo_camera_matrices.argtypes = [C.POINTER(Camera), _f32p, _f32p]
        L.yo_bound_sphere.argtypes = [C.c_float, _f32p, _f32p]
        L.yo_bound_triangle.argtypes = [_f32p] * 4
        L.yo_bb_from_points.argtypes = [_f32p] * 3
        L.yo_bb_include_point.argtypes = [_f32p] * 3
        L.yo_bb_join.argtypes = [_f32p] * 3
        L.yo_bb_centroid.argtypes = [_f32p] * 2
        L.yo_bb_surf.restype = C.c_float
        L.yo_bb_surf.argtypes = [_f32p]
        L.yo_bb_ray.restype = C.c_int
        L.yo_bb_ray.argtypes = [_f32p, _f32p, _f32p, C.c_float, _f32p]
        L.yo_collide_triangle.restype = C.c_int
        L.yo_collide_triangle.argtypes = [_f32p] * 8 + [C.c_float, _f32p]
        L.yo_collide_sphere.restype = C.c_int
        L.yo_collide_sphere.argtypes = [C.c_float, _f32p, _f32p, _f32p, C.c_float, _f32p]
        L.yo_scene_collide.restype = C.c_int
        L.yo_scene_collide.argtypes = [C.c_void_p, _f32p, _f32p, C.c_float, _f32p]
        L.yo_scene_radiance.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, C.c_float, _f32p]
        L.yo_bsdf_at.argtypes = [C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p]
        L.yo_sample_offset.restype = C.c_float
        L.yo_sample_offset.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.yo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def f32(a):
    """float32 C-contiguous array + pointer (keeps the array alive through the tuple)."""
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return arr, arr.ctypes.data_as(_f32p)


def _get(scene, name, default=None):
    if isinstance(scene, dict):
        return scene.get(name, default)
    return getattr(scene, name, default)


def scene_to_desc(scene, desc_cls=SceneDesc):
    """Build a ctypes scene descriptor; returns (desc, keepalive list)."""
    keep = []

    def fp(name, cols):
        a = _get(scene, name)
        if a is None or len(a) == 0:
            return None, 0
        arr = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, cols) if cols else
                                   np.asarray(a, dtype=np.float32).reshape(-1))
        keep.append(arr)
        return arr.ctypes.data_as(_f32p), arr.shape[0]

    def up(name, n):
        a = _get(scene, name)
        if a is None:
            return None
        arr = np.ascontiguousarray(np.asarray(a, dtype=np.uint32).reshape(-1))
        assert arr.shape[0] == n, (name, arr.shape, n)
        keep.append(arr)
        return arr.ctypes.data_as(_u32p)

    d = desc_cls()
    d.tri_p0, nt = fp("tri_p0", 3)
    for nme in ("tri_p1", "tri_p2", "tri_n0", "tri_n1", "tri_n2"):
        p, n = fp(nme, 3)
        assert n == nt, nme
        setattr(d, nme, p)
    d.n_triangles = nt
    d.tri_material = up("tri_material", nt) if nt else None
    d.sph_center, ns = fp("sph_center", 3)
    d.sph_radius, nr = fp("sph_radius", 0)
    assert nr == ns
    d.n_spheres = ns
    d.sph_material = up("sph_material", ns) if ns else None
    d.prim_order = up("prim_order", nt + ns) if _get(scene, "prim_order") is not None else None
    d.materials, d.n_materials = fp("materials", 7)
    d.lights, d.n_lights = fp("lights", 6)
    d.bvh_max_depth = int(_get(scene, "bvh_max_depth", 16))
    d.split_mode = int(_get(scene, "split_mode", 0))
    d.area_lights, d.n_area_lights = fp("area_lights", 13)      # extension: quad area lights
    return d, keep


def make_camera(cam, cam_cls=Camera):
    """cam: dict/object with imW, imH, focalLength, lookDir, upDir, position."""
    c = cam_cls()
    c.imW, c.imH, c.focalLength = float(_get(cam, "imW")), float(_get(cam, "imH")), float(_get(cam, "focalLength"))
    for n in ("lookDir", "upDir", "position"):
        v = _get(cam, n)
        getattr(c, n)[:] = [float(v[0]), float(v[1]), float(v[2])]
    return c


class OracleScene:
    """A built oracle scene (BVH built with the restated Culling.hs algorithm)."""

    def __init__(self, scene):
        self._L = lib()
        desc, self._keep = scene_to_desc(scene)
        self._h = self._L.yo_scene_create(C.byref(desc))
        if not self._h:
            raise RuntimeError("yo_scene_create failed")

    def close(self):
        if self._h:
            self._L.yo_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, cam, recursion_depth=1, spp=1, seed=0, threads=0, tile_stride=1, tile_offset=0,
               n_batches=0, out=None):
        """Returns (rgb[H,W,3] f32, primid[H,W] u32, t[H,W] f32, stats dict)."""
        c = make_camera(cam)
        w, h = int(np.floor(c.imW)), int(np.floor(c.imH))
        if out is None:
            rgb = np.zeros((h, w, 3), np.float32)
            pid = np.full((h, w), 0xFFFFFFFF, np.uint32)
            tt = np.zeros((h, w), np.float32)
        else:
            rgb, pid, tt = out
        st = Stats()
        rc = self._L.yo_render(self._h, C.byref(c), recursion_depth, spp, seed,
                               rgb.ctypes.data_as(_f32p), pid.ctypes.data_as(_u32p), tt.ctypes.data_as(_f32p),
                               C.byref(st), threads, tile_stride, tile_offset, n_batches)
        if rc != 0:
            raise RuntimeError("yo_render failed rc=%d" % rc)
        return rgb, pid, tt, st.as_dict()

    def collide(self, x0, u, tmax=1e6):
        a, pa = f32(x0)
        b, pb = f32(u)
        out = np.zeros(13, np.float32)
        pid = self._L.yo_scene_collide(self._h, pa, pb, tmax, out.ctypes.data_as(_f32p))
        return pid, out

    def radiance(self, x0, u, depth=1, tmax=1e6):
        a, pa = f32(x0)
        b, pb = f32(u)
        out = np.zeros(3, np.float32)
        self._L.yo_scene_radiance(self._h, depth, pa, pb, tmax, out.ctypes.data_as(_f32p))
        return out

    def bvh_order(self):
        n = self._L.yo_bvh_num_prims(self._h)
        out = np.zeros(n, np.uint32)
        self._L.yo_bvh_order(self._h, out.ctypes.data_as(_u32p))
        return out

    def bvh_depth(self):
        return self._L.yo_bvh_depth(self._h)

    def bvh_preorder(self):
        n = self._L.yo_bvh_num_nodes(self._h)
        kinds = np.zeros(n, np.uint32)
        firsts = np.zeros(n, np.uint32)
        counts = np.zeros(n, np.uint32)
        boxes = np.zeros((n, 6), np.float32)
        k = self._L.yo_bvh_dump_preorder(self._h, kinds.ctypes.data_as(_u32p), firsts.ctypes.data_as(_u32p),
                                         counts.ctypes.data_as(_u32p), boxes.ctypes.data_as(_f32p))
        return kinds[:k], firsts[:k], counts[:k], boxes[:k]


def bytes_per_ray(stats):
    """Algorithmic bytes per ray: 32 B per box test, 36 B per triangle test,
    36 B per triangle candidate (normals), 16 B per sphere test, 28 B per material fetch,
    24 B per light fetch, 12 B per primary ray's pixel write."""
    rays = stats["n_primary"] + stats["n_shadow"] + stats["n_secondary"]
    total = (32 * stats["n_node"] + 36 * stats["n_tri"] + 36 * stats["n_tricand"] + 16 * stats["n_sph"] +
             28 * stats["n_shade"] + 24 * stats["n_light"] + 12 * stats["n_primary"])
    return total / max(rays, 1), rays, total
