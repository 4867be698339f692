"""CPU tests of the product's host logic against the oracle: BVH construction (topology, DFS
order, boxes), camera matrices, tile arithmetic, and the C-ABI surface.  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import binding as ob
from yahr_b200 import api, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    api.build_library()


def small_scenes():
    out = {}
    out["c1"] = scenes.c1_scene_yahrr()[0]
    out["bunny"] = scenes.c2_bunny_proxy(64, 36, nu=40, nv=20)[0]
    out["grid"] = scenes.c3_sphere_grid(8, 64, 64)[0]
    out["terrain"] = scenes.c4_terrain(61, 31, 64, 36)[0]
    out["soup"] = scenes.c4_soup(5000, 64, 36)[0]
    out["adversarial"] = scenes.adversarial_shared_edges(32, 32)[0]
    # depth-capped variants produce multi-leaves
    capped = dict(out["bunny"])
    capped["bvh_max_depth"] = 5
    out["bunny-depth5"] = capped
    flat = dict(out["grid"])
    flat["bvh_max_depth"] = 0
    out["grid-depth0"] = flat
    return out


@pytest.mark.parametrize("split_mode", [0, 1])
@pytest.mark.parametrize("name", list(small_scenes().keys()))
def test_host_bvh_equals_oracle_tree(name, split_mode):
    sc = dict(small_scenes()[name])
    sc["split_mode"] = split_mode
    o = ob.OracleScene(sc)
    h = api.HostBvh(sc)
    assert np.array_equal(h.order(), o.bvh_order()), "DFS leaf order differs"
    ok, of, oc, obx = o.bvh_preorder()
    hk, hf, hc, hbx = h.preorder()
    assert np.array_equal(hk, ok), "node kinds (pre-order) differ"
    leaves = ok != 2
    assert np.array_equal(hf[leaves], of[leaves])
    assert np.array_equal(hc[leaves], oc[leaves])
    # boxes bit-exact (compare raw bits so that +-inf of Null nodes compare equal too)
    assert np.array_equal(hbx.view(np.uint32), obx.view(np.uint32))
    assert h.depth() == o.bvh_depth()
    o.close()
    h.close()


def test_host_bvh_empty_scene():
    sc = scenes._empty_scene()
    h = api.HostBvh(sc)
    assert len(h.order()) == 0
    h.close()


def test_unknown_material_is_an_error():
    sc = scenes.c1_scene_yahrr()[0]
    sc = dict(sc)
    sc["sph_material"] = np.full(7, 99, np.uint32)
    with pytest.raises(api.YahrError) as e:
        api.HostBvh(sc)
    assert e.value.code == 5


def test_non_finite_geometry_is_an_error():
    sc = dict(scenes.c1_scene_yahrr()[0])
    c = sc["sph_center"].copy()
    c[0, 0] = np.nan
    sc["sph_center"] = c
    with pytest.raises(api.YahrError) as e:
        api.HostBvh(sc)
    assert e.value.code == 7


@pytest.mark.parametrize("cam", [
    dict(imW=200, imH=200, focalLength=1, lookDir=[0, 0, 1], upDir=[0, 1, 0], position=[10, -80, 124]),
    dict(imW=1024, imH=768, focalLength=1.5, lookDir=[0.4, -0.3, 1], upDir=[0, 1, 0], position=[-4, 3, 2]),
    dict(imW=3840, imH=2160, focalLength=1.2, lookDir=[0, -0.42, 1], upDir=[0, 1, 0], position=[0, 38, -95]),
])
def test_camera_matrices_bit_equal_oracle(cam):
    tf, vtf = api.camera_matrices(cam)
    L = ob.lib()
    otf = np.zeros(16, np.float32)
    ovtf = np.zeros(16, np.float32)
    c = ob.make_camera(cam)
    L.yo_camera_matrices(C.byref(c), otf.ctypes.data_as(C.POINTER(C.c_float)), ovtf.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.array_equal(tf.reshape(-1).view(np.uint32), otf.view(np.uint32))
    assert np.array_equal(vtf.reshape(-1).view(np.uint32), ovtf.view(np.uint32))


@pytest.mark.parametrize("w,h,n", [(40, 40, 5), (40, 40, 16), (1, 1, 1), (20, 1, 3), (1, 20, 3), (3840, 2160, 32768),
                                   (1920, 1080, 8192)])
def test_tiles_equal_oracle_and_cover(w, h, n):
    L = ob.lib()
    seen = np.zeros((h, w), np.int32)
    win = (C.c_int32 * 4)()
    for i in range(n):
        L.yo_batch_window(w, h, i, n, win)
        assert api.batch_window(w, h, i, n) == tuple(win)
        x0, y0, x1, y1 = win
        seen[y0:y1, x0:x1] += 1
    assert (seen == 1).all()


@pytest.mark.parametrize("t,w,h", [(1, 512, 512), (8, 1920, 1080), (8, 2048, 2048), (8, 3840, 2160), (1, 10, 10),
                                   (4, 100, 3)])
def test_num_batches_equal_oracle(t, w, h):
    assert api.num_batches(t, w, h) == ob.lib().yo_num_batches(t, w, h)


def test_abi_exports_every_declared_symbol():
    """The shared library loads and exports every function include/yahr_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "yahr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(yahr_b200_\w+)\s*\(", hdr))
    assert len(names) >= 18
    L = C.CDLL(api.LIB_PATH)
    for n in sorted(names):
        assert hasattr(L, n), "missing export " + n
    assert api.lib().yahr_b200_abi_version() == 2


def test_binding_struct_sizes_match_the_library():
    """The ctypes mirrors in yahr_b200/api.py (and the byte offsets poked by hs/GpuRender.hs: 160-byte scene descriptor)
    must have the sizes the C library was compiled with."""
    L = api.lib()
    L.yahr_b200_sizeof.restype = C.c_int
    for which, cls in enumerate([api.SceneDesc, api.Camera, api.RenderOpts, api.Stats, api.SceneInfo]):
        assert L.yahr_b200_sizeof(which) == C.sizeof(cls), cls.__name__
    assert L.yahr_b200_sizeof(0) == 160
    assert "allocaBytes 160" in open(os.path.join(ROOT, "hs", "GpuRender.hs")).read()
    assert L.yahr_b200_sizeof(99) == -1


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device scene_create must fail with YAHR_ERR_NO_DEVICE (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.YahrError) as e:
        api.Scene(scenes.c1_scene_yahrr()[0])
    assert e.value.code in (2, 3)


def test_product_does_not_reference_the_oracle():
    """The product path must not import, link or include anything under oracle/."""
    bad = []
    for base in ("yahr_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|yahr_oracle|liboracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_out_buffer_validation_is_strict():
    """api._check_out guards every raw pointer handed to the C ABI: dtype, exact shape, C-contiguity, writeability."""
    import numpy as np
    api._check_out(np.zeros((4, 6, 3), np.float32), (4, 6, 3), np.float32, "rgb")
    bad = [np.zeros((4, 6, 3), np.float64), np.zeros((6, 4, 3), np.float32), np.zeros((4, 12, 3), np.float32)[:, ::2],
           np.zeros((4, 6, 3), np.float32).T, [[0.0] * 18] * 4]
    ro = np.zeros((4, 6, 3), np.float32)
    ro.setflags(write=False)
    for a in bad + [ro]:
        with pytest.raises(ValueError):
            api._check_out(a, (4, 6, 3), np.float32, "rgb")


def test_bench_refuses_a_stale_ncu_capture(monkeypatch):
    """bench.py takes its per-launch instruction / wavefront counts from a committed ncu summary only when that summary
    was taken on the sources the loaded library is built from (fingerprint of csrc/* + the header)."""
    import json
    import bench
    path = os.path.join(ROOT, bench.NCU_SUMMARIES["c4-terrain"])
    js = json.load(open(path))
    assert len(js["fingerprint"]) == 16 and any("k_wf_primary" in k["name"] for k in js["kernels"])
    monkeypatch.setattr(api, "source_fingerprint", lambda: js["fingerprint"])
    cap, why = bench.ncu_capture("c4-terrain")
    assert why is None and cap["kernels"]["k_wf_primary"]["inst_executed"] > 1e8
    monkeypatch.setattr(api, "source_fingerprint", lambda: "0" * 16)
    cap, why = bench.ncu_capture("c4-terrain")
    assert cap is None and "refused" in why
    cap, why = bench.ncu_capture("no-such-workload")
    assert cap is None and why


def test_source_fingerprint_tracks_the_kernel_sources():
    fp = api.source_fingerprint()
    assert len(fp) == 16 and fp == api.source_fingerprint()
