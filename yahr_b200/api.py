"""ctypes binding of libyahr_b200.so (include/yahr_b200.h).

This is plumbing: every render goes through the C ABI into the CUDA kernels.  There is no CPU
fallback -- a missing library or a missing GPU raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# YAHR_B200_LIB: another build of the same library (experiments with compile-time variants); default: the in-tree build
LIB_PATH = os.environ.get("YAHR_B200_LIB") or os.path.join(_HERE, "libyahr_b200.so")
CSRC = os.path.join(_HERE, "csrc")

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)

TRAVERSAL_REFERENCE = 0
TRAVERSAL_ORDERED = 1
PRIM_MISS = 0xFFFFFFFF


class YahrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("yahr_b200 error %d: %s" % (code, msg))
        self.code = code


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


class RenderOpts(C.Structure):
    _fields_ = [("recursion_depth", C.c_int32), ("spp", C.c_int32), ("seed", C.c_uint64),
                ("traversal", C.c_int32), ("tile_stride", C.c_int32), ("tile_offset", C.c_int32),
                ("kernel", C.c_int32), ("reserved", C.c_int32 * 4)]


class Stats(C.Structure):
    _fields_ = [("n_primary", C.c_uint64), ("n_shadow", C.c_uint64), ("n_secondary", C.c_uint64),
                ("gpu_ms", C.c_double), ("wall_ms", C.c_double), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("launches", C.c_uint32), ("tiles", C.c_uint32),
                ("phase_ms", C.c_double * 4)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["phase_ms"] = list(self.phase_ms)
        return d


class SceneInfo(C.Structure):
    _fields_ = [("n_primitives", C.c_uint32), ("n_nodes", C.c_uint32), ("n_multi_leaves", C.c_uint32),
                ("depth", C.c_uint32), ("device_bytes", C.c_uint64), ("build_ms", C.c_double),
                ("upload_ms", C.c_double), ("built_on_device", C.c_uint32), ("n_wide_nodes", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build_library(force=False):
    """Compile libyahr_b200.so in-tree with nvcc for sm_100a (no GPU needed to build)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
            if f.endswith((".cu", ".cpp", ".hpp", ".cuh")) or f == "Makefile"]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "yahr_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if os.environ.get("YAHR_B200_LIB"):
        return LIB_PATH                      # a hand-built variant: never rebuilt behind the caller's back
    if force or stale:
        subprocess.check_call(["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", CSRC], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library.  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libyahr_b200.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; "
                          "g.build()'` or `make -C yahr_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.yahr_b200_abi_version.restype = C.c_int
    L.yahr_b200_device_count.restype = C.c_int
    L.yahr_b200_last_error.restype = C.c_char_p
    L.yahr_b200_scene_create.restype = C.c_int
    L.yahr_b200_scene_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]
    L.yahr_b200_scene_destroy.argtypes = [C.c_void_p]
    L.yahr_b200_scene_info.restype = C.c_int
    L.yahr_b200_scene_info.argtypes = [C.c_void_p, C.POINTER(SceneInfo)]
    L.yahr_b200_render.restype = C.c_int
    L.yahr_b200_render.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_int, C.c_int, C.c_uint64, C.c_void_p,
                                   C.c_void_p, C.POINTER(Stats)]
    L.yahr_b200_render_device.restype = C.c_int
    L.yahr_b200_render_device.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(RenderOpts), C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.POINTER(Stats)]
    L.yahr_b200_num_batches.restype = C.c_int64
    L.yahr_b200_num_batches.argtypes = [C.c_int64] * 3
    L.yahr_b200_batch_window.restype = C.c_int
    L.yahr_b200_batch_window.argtypes = [C.c_int64] * 4 + [C.POINTER(C.c_int32)]
    L.yahr_b200_ipc_export.restype = C.c_int
    L.yahr_b200_ipc_export.argtypes = [C.c_void_p, C.c_char_p]
    L.yahr_b200_ipc_open.restype = C.c_int
    L.yahr_b200_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.yahr_b200_ipc_close.restype = C.c_int
    L.yahr_b200_ipc_close.argtypes = [C.c_void_p]
    L.yahr_b200_host_bvh_build.restype = C.c_int
    L.yahr_b200_host_bvh_build.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]
    L.yahr_b200_host_bvh_destroy.argtypes = [C.c_void_p]
    for n in ("num_primitives", "num_nodes", "depth"):
        f = getattr(L, "yahr_b200_host_bvh_" + n)
        f.restype = C.c_uint32
        f.argtypes = [C.c_void_p]
    L.yahr_b200_host_bvh_order.restype = C.c_int
    L.yahr_b200_host_bvh_order.argtypes = [C.c_void_p, _u32p]
    L.yahr_b200_host_bvh_preorder.restype = C.c_int
    L.yahr_b200_host_bvh_preorder.argtypes = [C.c_void_p, _u32p, _u32p, _u32p, _f32p]
    L.yahr_b200_camera_matrices.restype = C.c_int
    L.yahr_b200_camera_matrices.argtypes = [C.POINTER(Camera), _f32p, _f32p]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise YahrError(rc, lib().yahr_b200_last_error().decode("utf-8", "replace"))


def _get(obj, name, default=None):
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


def scene_desc(scene):
    """dict/object of numpy arrays -> (SceneDesc, keepalive)."""
    keep = []

    def fp(name, cols):
        a = _get(scene, name)
        if a is None or len(a) == 0:
            return None, 0
        arr = np.asarray(a, dtype=np.float32)
        arr = np.ascontiguousarray(arr.reshape(-1, cols) if cols else arr.reshape(-1))
        keep.append(arr)
        return arr.ctypes.data_as(_f32p), arr.shape[0]

    def up(name, n):
        a = _get(scene, name)
        if a is None or n == 0:
            return None
        arr = np.ascontiguousarray(np.asarray(a, dtype=np.uint32).reshape(-1))
        if arr.shape[0] != n:
            raise ValueError("%s has %d entries, expected %d" % (name, arr.shape[0], n))
        keep.append(arr)
        return arr.ctypes.data_as(_u32p)

    d = SceneDesc()
    d.tri_p0, nt = fp("tri_p0", 3)
    for nme in ("tri_p1", "tri_p2", "tri_n0", "tri_n1", "tri_n2"):
        p, n = fp(nme, 3)
        if n != nt:
            raise ValueError("%s has %d rows, expected %d" % (nme, n, nt))
        setattr(d, nme, p)
    d.n_triangles = nt
    d.tri_material = up("tri_material", nt)
    d.sph_center, ns = fp("sph_center", 3)
    d.sph_radius, nr = fp("sph_radius", 0)
    if nr != ns:
        raise ValueError("sph_radius has %d entries, expected %d" % (nr, ns))
    d.n_spheres = ns
    d.sph_material = up("sph_material", ns)
    d.prim_order = up("prim_order", nt + ns) if _get(scene, "prim_order") is not None else None
    d.materials, d.n_materials = fp("materials", 7)
    d.lights, d.n_lights = fp("lights", 6)
    d.bvh_max_depth = int(_get(scene, "bvh_max_depth", 16))
    d.split_mode = int(_get(scene, "split_mode", 0))
    d.area_lights, d.n_area_lights = fp("area_lights", 13)      # extension: quad area lights
    return d, keep


def make_camera(cam):
    c = Camera()
    c.imW, c.imH, c.focalLength = float(_get(cam, "imW")), float(_get(cam, "imH")), float(_get(cam, "focalLength"))
    for n in ("lookDir", "upDir", "position"):
        v = _get(cam, n)
        getattr(c, n)[:] = [float(v[0]), float(v[1]), float(v[2])]
    return c


def image_size(cam):
    c = make_camera(cam)
    return int(np.floor(c.imW)), int(np.floor(c.imH))


def source_fingerprint():
    """sha256 over the sources libyahr_b200.so is built from (csrc/*, include/yahr_b200.h).  Profiles written by
    tools/ncu_summary.py carry it, and bench.py only trusts a capture whose fingerprint is the current tree's."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", ".cpp")) or f == "Makefile")
    for f in files:
        h.update(f.encode())
        h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(os.path.dirname(_HERE), "include", "yahr_b200.h"), "rb").read())
    return h.hexdigest()[:16]


def _check_out(a, shape, dtype, what):
    """The C ABI writes shape-many elements of dtype through a raw pointer: refuse anything that is not exactly that."""
    if not isinstance(a, np.ndarray):
        raise ValueError("%s must be a numpy array" % what)
    if a.dtype != np.dtype(dtype) or tuple(a.shape) != tuple(shape) or not a.flags.c_contiguous or not a.flags.writeable:
        raise ValueError("%s must be a writeable C-contiguous %s array of shape %s (got %s %s%s)"
                         % (what, np.dtype(dtype).name, tuple(shape), a.dtype.name, tuple(a.shape),
                            "" if a.flags.c_contiguous else ", not contiguous"))


def device_count():
    return lib().yahr_b200_device_count()


class Scene:
    """A scene resident on the current CUDA device (BVH built and uploaded once)."""

    def __init__(self, scene):
        L = lib()
        d, keep = scene_desc(scene)
        h = C.c_void_p()
        _check(L.yahr_b200_scene_create(C.byref(d), C.byref(h)))
        self._h = h
        del keep

    def close(self):
        if getattr(self, "_h", None):
            lib().yahr_b200_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        i = SceneInfo()
        _check(lib().yahr_b200_scene_info(self._h, C.byref(i)))
        return i.as_dict()

    def render(self, cam, recursion_depth=1, spp=1, seed=0, want_primid=True, out=None):
        """Host-buffer entry (yahr_b200_render).  Returns (rgb[H,W,3], primid[H,W] or None, stats)."""
        c = make_camera(cam)
        w, h = int(np.floor(c.imW)), int(np.floor(c.imH))
        if out is None:
            rgb = np.empty((h, w, 3), np.float32)
            pid = np.empty((h, w), np.uint32) if want_primid else None
        else:
            rgb, pid = out
            _check_out(rgb, (h, w, 3), np.float32, "out[0] (rgb)")
            if pid is not None:
                _check_out(pid, (h, w), np.uint32, "out[1] (primid)")
        st = Stats()
        _check(lib().yahr_b200_render(self._h, C.byref(c), recursion_depth, spp, seed, rgb.ctypes.data,
                                      pid.ctypes.data if pid is not None else None, C.byref(st)))
        return rgb, pid, st.as_dict()

    def render_shard(self, cam, shard_index, shard_count, out, recursion_depth=1, spp=1, seed=0):
        """yahr_b200_render_shard: renders the tile rows shard_index, shard_index + shard_count, ... into the
        full-frame host buffers out = (rgb [H,W,3] float32, primid [H,W] uint32 or None)."""
        c = make_camera(cam)
        rgb, pid = out
        w, h = int(np.floor(c.imW)), int(np.floor(c.imH))
        if hasattr(rgb, "ctypes"):
            _check_out(rgb, (h, w, 3), np.float32, "out[0] (rgb)")
        if pid is not None and hasattr(pid, "ctypes"):
            _check_out(pid, (h, w), np.uint32, "out[1] (primid)")
        st = Stats()
        L = lib()
        L.yahr_b200_render_shard.restype = C.c_int
        L.yahr_b200_render_shard.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        _check(L.yahr_b200_render_shard(self._h, C.byref(c), recursion_depth, spp, seed, shard_index, shard_count,
                                        rgb.ctypes.data if hasattr(rgb, "ctypes") else int(rgb),
                                        (pid.ctypes.data if hasattr(pid, "ctypes") else int(pid)) if pid is not None else None,
                                        C.byref(st)))
        return st.as_dict()

    def render_shard_rgb8(self, cam, shard_index, shard_count, out, recursion_depth=1, spp=1, seed=0):
        """yahr_b200_render_shard_rgb8: this shard's tile rows of the 8-bit frame into the full-frame host buffer
        out ([H,W,3] uint8, or a raw address)."""
        c = make_camera(cam)
        w, h = int(np.floor(c.imW)), int(np.floor(c.imH))
        if hasattr(out, "ctypes"):
            _check_out(out, (h, w, 3), np.uint8, "out (rgb8)")
        st = Stats()
        L = lib()
        L.yahr_b200_render_shard_rgb8.restype = C.c_int
        L.yahr_b200_render_shard_rgb8.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_int, C.c_int, C.c_uint64, C.c_int,
                                                  C.c_int, C.c_void_p, C.POINTER(Stats)]
        _check(L.yahr_b200_render_shard_rgb8(self._h, C.byref(c), recursion_depth, spp, seed, shard_index, shard_count,
                                             out.ctypes.data if hasattr(out, "ctypes") else int(out), C.byref(st)))
        return st.as_dict()

    def download_bvh(self):
        """(order[n_prims], nodes[n_nodes,16] float32 raw, multi[n_multi,2], root_ref, root_box[6])."""
        i = self.info()
        order = np.zeros(i["n_primitives"], np.uint32)
        nodes = np.zeros((i["n_nodes"], 16), np.float32)
        multi = np.zeros((i["n_multi_leaves"], 2), np.uint32)
        root = C.c_uint32(0)
        box = np.zeros(6, np.float32)
        L = lib()
        L.yahr_b200_scene_download_bvh.restype = C.c_int
        L.yahr_b200_scene_download_bvh.argtypes = [C.c_void_p] * 4 + [C.POINTER(C.c_uint32), C.c_void_p]
        _check(L.yahr_b200_scene_download_bvh(self._h, order.ctypes.data, nodes.ctypes.data, multi.ctypes.data,
                                              C.byref(root), box.ctypes.data))
        return order, nodes, multi, root.value, box

    def download_wide(self):
        """wide[n_wide_nodes, 32] float32 raw: the 4-wide collapse of the tree (csrc/device_types.cuh)."""
        i = self.info()
        wide = np.zeros((i["n_wide_nodes"], 32), np.float32)
        L = lib()
        L.yahr_b200_scene_download_wide.restype = C.c_int
        L.yahr_b200_scene_download_wide.argtypes = [C.c_void_p, C.c_void_p]
        _check(L.yahr_b200_scene_download_wide(self._h, wide.ctypes.data))
        return wide

    def download_compressed(self):
        """(cwide[n_wide, 16] uint32 raw, leaf_box[n_prims, 8] float32, multi_box[n_multi, 8] float32): the compressed
        wide nodes and the exact leaf boxes of the compressed walk (csrc/wide_bvh.cu)."""
        i = self.info()
        cw = np.zeros((i["n_wide_nodes"], 16), np.uint32)
        lb = np.zeros((i["n_primitives"], 8), np.float32)
        mb = np.zeros((i["n_multi_leaves"], 8), np.float32)
        L = lib()
        L.yahr_b200_scene_download_compressed.restype = C.c_int
        L.yahr_b200_scene_download_compressed.argtypes = [C.c_void_p] * 4
        _check(L.yahr_b200_scene_download_compressed(self._h, cw.ctypes.data, lb.ctypes.data, mb.ctypes.data))
        return cw, lb, mb

    def render_rgb8(self, cam, recursion_depth=1, spp=1, seed=0, out=None):
        """yahr_b200_render_rgb8: the frame with the reference's 8-bit output stage applied on the GPU."""
        c = make_camera(cam)
        w, h = int(np.floor(c.imW)), int(np.floor(c.imH))
        rgb8 = np.empty((h, w, 3), np.uint8) if out is None else out
        _check_out(rgb8, (h, w, 3), np.uint8, "out (rgb8)")
        L = lib()
        L.yahr_b200_render_rgb8.restype = C.c_int
        L.yahr_b200_render_rgb8.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_int, C.c_int, C.c_uint64, C.c_void_p,
                                            C.POINTER(Stats)]
        st = Stats()
        _check(L.yahr_b200_render_rgb8(self._h, C.byref(c), recursion_depth, spp, seed, rgb8.ctypes.data, C.byref(st)))
        return rgb8, st.as_dict()

    def render_device(self, cam, d_rgb, d_primid=None, recursion_depth=1, spp=1, seed=0,
                      traversal=TRAVERSAL_REFERENCE, tile_stride=1, tile_offset=0, stream=None, stats=True,
                      kernel=0, tune=0, by_rows=False):
        """Device-buffer entry (yahr_b200_render_device).  d_rgb / d_primid are raw device pointers
        (ints), e.g. torch_tensor.data_ptr().  Returns a stats dict (synchronises) or None."""
        c = make_camera(cam)
        o = RenderOpts()
        o.recursion_depth, o.spp, o.seed = recursion_depth, spp, seed
        o.traversal, o.tile_stride, o.tile_offset, o.kernel = traversal, tile_stride, tile_offset, kernel
        o.reserved[0] = tune
        o.reserved[1] = 1 if by_rows else 0     # tile_stride / tile_offset count whole rows of the tile grid
        st = Stats() if stats else None
        _check(lib().yahr_b200_render_device(self._h, C.byref(c), C.byref(o), C.c_void_p(d_rgb),
                                             C.c_void_p(d_primid) if d_primid else None,
                                             C.c_void_p(stream) if stream else None,
                                             C.byref(st) if stats else None))
        return st.as_dict() if stats else None


def _ptr(x):
    return C.c_void_p(x) if x else None


def _render_device_counted(self, cam, d_rgb, d_primid=None, recursion_depth=1, spp=1, seed=0,
                           traversal=TRAVERSAL_REFERENCE, tile_stride=1, tile_offset=0, stream=None, kernel=0, tune=0,
                           by_rows=False):
    """yahr_b200_render_device_counted: the frame through the counting build of the wavefront kernels.  Returns a
    dict of the GPU's own work counters and the bytes the kernels request per ray (measurement aid)."""
    c = make_camera(cam)
    o = RenderOpts()
    o.recursion_depth, o.spp, o.seed = recursion_depth, spp, seed
    o.traversal, o.tile_stride, o.tile_offset, o.kernel = traversal, tile_stride, tile_offset, kernel
    o.reserved[0] = tune
    o.reserved[1] = 1 if by_rows else 0
    st = Stats()
    counts = (C.c_uint64 * 16)()
    L = lib()
    L.yahr_b200_render_device_counted.restype = C.c_int
    L.yahr_b200_render_device_counted.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(RenderOpts), C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.POINTER(Stats), C.POINTER(C.c_uint64)]
    _check(L.yahr_b200_render_device_counted(self._h, C.byref(c), C.byref(o), C.c_void_p(d_rgb), _ptr(d_primid),
                                             _ptr(stream), C.byref(st), counts))
    names = ("wide_nodes", "binary_nodes", "prim_tests", "normal_fetches", "stack_pushes", "stack_pops", "_", "shaded")
    out = {"closest_hit": {n: int(counts[k]) for k, n in enumerate(names) if n != "_"},
           "any_hit": {n: int(counts[8 + k]) for k, n in enumerate(names) if n not in ("_", "shaded")}}
    s_ = st.as_dict()
    rays = s_["n_primary"] + s_["n_shadow"] + s_["n_secondary"]
    tot = {n: out["closest_hit"].get(n, 0) + out["any_hit"].get(n, 0) for n in names if n != "_"}
    # bytes the kernels REQUEST (from L1/L2/HBM, whatever serves them): 128 B per 4-wide node, 64 B per binary node, 48 B
    # per primitive test, 48 B per normal fetch, 8 B per stack push / pop; per shaded hit the surface (48 + 48 B),
    # material (32 B) and light (32 B); per probe its queue record written and read; per primary ray the pixel-table
    # entry (4 B) and the pixel (12 B)
    probe_bytes = 2 * 32 if s_["launches"] <= 2 else 2 * 48      # k_wf_persist ring record / two-kernel queue record
    b = (128 * tot["wide_nodes"] + 64 * tot["binary_nodes"] + 48 * tot["prim_tests"] + 48 * tot["normal_fetches"]
         + 8 * (tot["stack_pushes"] + tot["stack_pops"]) + 160 * tot["shaded"] + probe_bytes * s_["n_shadow"]
         + 16 * s_["n_primary"])
    out.update({"rays": int(rays), "n_primary": int(s_["n_primary"]), "n_shadow": int(s_["n_shadow"]),
                "bytes_requested": int(b), "bytes_per_ray_gpu": b / max(rays, 1),
                "wide_nodes_per_ray": tot["wide_nodes"] / max(rays, 1), "prim_tests_per_ray": tot["prim_tests"] / max(rays, 1),
                "stack_bytes_share": 8 * (tot["stack_pushes"] + tot["stack_pops"]) / max(b, 1),
                "probe_record_bytes": probe_bytes // 2})
    return out


Scene.render_device_counted = _render_device_counted


def render_device_shard(scene, cam, shard_index, shard_count, d_rgb_local, d_rgb_gather=None, d_primid_local=None,
                        d_primid_gather=None, recursion_depth=1, spp=1, seed=0, traversal=TRAVERSAL_REFERENCE,
                        stream=None, stats=False, kernel=0, tune=0):
    """yahr_b200_render_device_shard: this shard's tile rows into the local buffers, then pushed into the gather
    buffers (raw device pointers, possibly peer memory) with device-to-device copies on `stream`."""
    c = make_camera(cam)
    o = RenderOpts()
    o.recursion_depth, o.spp, o.seed = recursion_depth, spp, seed
    o.traversal, o.tile_stride, o.tile_offset, o.kernel = traversal, 1, 0, kernel
    o.reserved[0] = tune
    st = Stats() if stats else None
    L = lib()
    L.yahr_b200_render_device_shard.restype = C.c_int
    L.yahr_b200_render_device_shard.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(RenderOpts), C.c_int, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.POINTER(Stats)]
    _check(L.yahr_b200_render_device_shard(scene._h, C.byref(c), C.byref(o), shard_index, shard_count, _ptr(d_rgb_local),
                                           _ptr(d_rgb_gather), _ptr(d_primid_local), _ptr(d_primid_gather), _ptr(stream),
                                           C.byref(st) if stats else None))
    return st.as_dict() if stats else None


class HostBvh:
    """Host-only view of the BVH the library would upload (no CUDA needed)."""

    def __init__(self, scene):
        d, keep = scene_desc(scene)
        h = C.c_void_p()
        _check(lib().yahr_b200_host_bvh_build(C.byref(d), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().yahr_b200_host_bvh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def depth(self):
        return lib().yahr_b200_host_bvh_depth(self._h)

    def order(self):
        n = lib().yahr_b200_host_bvh_num_primitives(self._h)
        out = np.zeros(n, np.uint32)
        _check(lib().yahr_b200_host_bvh_order(self._h, out.ctypes.data_as(_u32p)))
        return out

    def preorder(self):
        n = lib().yahr_b200_host_bvh_num_nodes(self._h)
        kinds, firsts, counts = (np.zeros(n, np.uint32) for _ in range(3))
        boxes = np.zeros((n, 6), np.float32)
        _check(lib().yahr_b200_host_bvh_preorder(self._h, kinds.ctypes.data_as(_u32p), firsts.ctypes.data_as(_u32p),
                                                 counts.ctypes.data_as(_u32p), boxes.ctypes.data_as(_f32p)))
        return kinds, firsts, counts, boxes


def load_yahrr(text):
    """Parse `.yahrr` text (derived-Read syntax of Scene, Scene.hs:52-58) and apply `expand`
    (Scene.hs:61-86) with the library's C++ host code.  Returns (scene dict, camera dict, recursionDepth)."""
    L = lib()
    L.yahr_b200_yahrr_load.restype = C.c_int
    L.yahr_b200_yahrr_load.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.yahr_b200_yahrr_free.argtypes = [C.c_void_p]
    L.yahr_b200_yahrr_describe.restype = C.c_int
    L.yahr_b200_yahrr_describe.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(Camera), C.POINTER(C.c_int)]
    data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    h = C.c_void_p()
    _check(L.yahr_b200_yahrr_load(data, len(data), C.byref(h)))
    try:
        d, c, depth = SceneDesc(), Camera(), C.c_int(0)
        _check(L.yahr_b200_yahrr_describe(h, C.byref(d), C.byref(c), C.byref(depth)))

        def arr(ptr, n, cols, dtype):
            if n == 0 or not ptr:
                return np.zeros((0, cols) if cols else (0,), dtype)
            a = np.ctypeslib.as_array(ptr, shape=(n * max(cols, 1),)).astype(dtype, copy=True)
            return a.reshape(n, cols) if cols else a

        nt, ns = d.n_triangles, d.n_spheres
        scene = dict(
            tri_p0=arr(d.tri_p0, nt, 3, np.float32), tri_p1=arr(d.tri_p1, nt, 3, np.float32),
            tri_p2=arr(d.tri_p2, nt, 3, np.float32), tri_n0=arr(d.tri_n0, nt, 3, np.float32),
            tri_n1=arr(d.tri_n1, nt, 3, np.float32), tri_n2=arr(d.tri_n2, nt, 3, np.float32),
            tri_material=arr(d.tri_material, nt, 0, np.uint32),
            sph_center=arr(d.sph_center, ns, 3, np.float32), sph_radius=arr(d.sph_radius, ns, 0, np.float32),
            sph_material=arr(d.sph_material, ns, 0, np.uint32),
            prim_order=arr(d.prim_order, nt + ns, 0, np.uint32),
            materials=arr(d.materials, d.n_materials, 7, np.float32), lights=arr(d.lights, d.n_lights, 6, np.float32),
            bvh_max_depth=int(d.bvh_max_depth), split_mode=int(d.split_mode))
        cam = dict(imW=float(c.imW), imH=float(c.imH), focalLength=float(c.focalLength),
                   lookDir=list(c.lookDir), upDir=list(c.upDir), position=list(c.position))
        return scene, cam, int(depth.value)
    finally:
        L.yahr_b200_yahrr_free(h)


def write_png_rgb8(path, rgb8):
    a = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w = a.shape[0], a.shape[1]
    L = lib()
    L.yahr_b200_write_png_rgb8.restype = C.c_int
    L.yahr_b200_write_png_rgb8.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    _check(L.yahr_b200_write_png_rgb8(path.encode(), a.ctypes.data, w, h))


def quantize_rgb8_host(rgb):
    a = np.ascontiguousarray(rgb, dtype=np.float32)
    out = np.empty(a.shape, np.uint8)
    L = lib()
    L.yahr_b200_quantize_rgb8_host.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.yahr_b200_quantize_rgb8_host(a.ctypes.data, a.size, out.ctypes.data)
    return out


def camera_matrices(cam):
    c = make_camera(cam)
    tf = np.zeros(16, np.float32)
    vtf = np.zeros(16, np.float32)
    _check(lib().yahr_b200_camera_matrices(C.byref(c), tf.ctypes.data_as(_f32p), vtf.ctypes.data_as(_f32p)))
    return tf.reshape(4, 4), vtf.reshape(4, 4)


def num_batches(num_threads, w, h):
    return lib().yahr_b200_num_batches(num_threads, w, h)


def batch_window(w, h, num, count):
    out = (C.c_int32 * 4)()
    _check(lib().yahr_b200_batch_window(w, h, num, count, out))
    return tuple(out)
