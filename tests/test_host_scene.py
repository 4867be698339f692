"""CPU tests of the host-side scene reader: the `.yahrr` grammar (derived Read of Scene.hs:15-58),
`Scene.expand` (Scene.hs:61-86; KAT: Spec.hs:148-172), the PNG output stage (main.hs:142) and the
golden fixtures written by the reference's own compat/yahr.py (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from yahr_b200 import api, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
F = np.float32


@pytest.fixture(scope="module", autouse=True)
def _built():
    api.build_library()


def wrap(objects, materials='BlinnPhongMaterial { id = "fooMaterial", ambient = Vec3 0 0 0, diffuse = Vec3 1 1 1, '
                            'specular = Vec3 0 0 0, shininess = 1 }'):
    return ('Scene { integrator = WhittedIntegrator { recursionDepth = 1 }, cullingMode = BVH 16 Midpoint, '
            'camera = Camera { imW = 8, imH = 8, focalLength = 1, lookDir = Vec3 0 0 1, upDir = Vec3 0 1 0, '
            'position = Vec3 0 0 0 }, materials = [ %s ], lights = [], objects = [ %s ] }' % (materials, objects))


def test_expand_keeps_vertex_order_and_material():
    """Spec.hs:148-172: a 4-point / 2-triangle mesh."""
    sc, cam, depth = api.load_yahrr(wrap(
        'TriangleMesh { triangleMeshPoints = [Vec3 0 0 0, Vec3 0 0 1, Vec3 0 1 0, Vec3 0 1 1], '
        'triangleMeshNormals = Nothing, triangleMeshTriangles = [(0, 1, 2), (1, 2, 3)], '
        'triangleMeshSmooth = Nothing, materialId = "fooMaterial" }'))
    assert sc["tri_p0"].tolist() == [[0, 0, 0], [0, 0, 1]]
    assert sc["tri_p1"].tolist() == [[0, 0, 1], [0, 1, 0]]
    assert sc["tri_p2"].tolist() == [[0, 1, 0], [0, 1, 1]]
    assert sc["tri_material"].tolist() == [0, 0]
    # flat normal n = norm ((p2 - p0) x (p1 - p0))  (Scene.hs:78), replicated on the three vertices
    n = scenes.flat_normals(sc["tri_p0"], sc["tri_p1"], sc["tri_p2"])
    for k in ("tri_n0", "tri_n1", "tri_n2"):
        assert np.array_equal(sc[k].view(np.uint32), n.view(np.uint32))
    assert (sc["prim_order"] >> 31).tolist() == [1, 1]


@pytest.mark.parametrize("name", ["testscene", "mixed"])
def test_golden_files_written_by_the_reference_scene_writer(name):
    text = open(os.path.join(GOLD, name + ".yahr")).read()
    exp = json.load(open(os.path.join(GOLD, name + ".expected.json")))
    sc, cam, depth = api.load_yahrr(text)
    assert depth == exp["recursionDepth"]
    assert sc["bvh_max_depth"] == exp["bvhMaxDepth"]
    assert sc["split_mode"] == (1 if exp["splitMode"] == "SurfaceAreaHeuristic" else 0)
    for k in ("imW", "imH", "focalLength"):
        assert cam[k] == exp["camera"][k]
    for k in ("lookDir", "upDir", "position"):
        assert np.array_equal(np.asarray(cam[k], F), np.asarray(exp["camera"][k], F))
    assert len(sc["materials"]) == len(exp["materials"])
    for row, m in zip(sc["materials"], exp["materials"]):
        assert np.array_equal(row, np.asarray(m["diffuse"] + m["specular"] + [m["shininess"]], F))
    assert np.array_equal(sc["lights"], np.asarray(exp["lights"], F).reshape(-1, 6))
    prims = exp["prims"]
    assert len(sc["tri_p0"]) == len(prims)
    for i, p in enumerate(prims):
        for k, nm in enumerate(("tri_p0", "tri_p1", "tri_p2")):
            assert np.array_equal(sc[nm][i], np.asarray(p["p"][k], F)), (i, nm)
        for k, nm in enumerate(("tri_n0", "tri_n1", "tri_n2")):
            a, b = sc[nm][i], np.asarray(p["n"][k], F)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), (i, nm)
        assert sc["tri_material"][i] == p["material"]       # "dup": the later duplicate id wins (Map.fromList)


def test_repo_scene_in_current_schema_equals_generator():
    sc, cam, depth = api.load_yahrr(open(os.path.join(ROOT, "scenes", "scene.yahrr")).read())
    ref, rcam = scenes.c1_scene_yahrr(recursion_native=True)
    assert depth == 3 and (cam["imW"], cam["imH"]) == (1024.0, 768.0)
    for k in ("sph_center", "sph_radius", "tri_p0", "tri_p1", "tri_p2", "tri_n0", "materials", "lights"):
        assert np.array_equal(np.asarray(sc[k], F).reshape(-1), np.asarray(ref[k], F).reshape(-1)), k
    assert sc["sph_material"].tolist() == [0] * 7 and sc["tri_material"].tolist() == [3, 3]
    # spheres come first in the object list, then the floor
    assert (sc["prim_order"] >> 31).tolist() == [0] * 7 + [1, 1]


def test_stale_syntax_leniency():
    """The repo's own scene.yahrr omits the two Maybe fields of TriangleMesh and writes bare negative
    numbers (scene.yahrr:6,35,56-65); both are accepted, the missing fields default to Nothing."""
    sc, _, _ = api.load_yahrr(wrap(
        'TriangleMesh { triangleMeshPoints = [ Vec3 -1000 -2.2 -1000, Vec3 1000 -2.2 -1000, Vec3 1000 -2.2 1000 ], '
        'triangleMeshTriangles = [ (0, 1, 2) ], materialId = "fooMaterial" }'))
    assert sc["tri_p0"].tolist() == [[-1000, F(-2.2), -1000]]
    assert np.array_equal(sc["tri_n0"], sc["tri_n1"])


def test_redundant_parentheses_numbers_and_comments():
    sc, cam, _ = api.load_yahrr(
        "-- a comment\n(" + wrap('(Sphere { position = ((Vec3 (1) (-2.5e0) 3e-1)), radius = (2), materialId = ("fooMaterial") }), '
                               'Triangle { p0 = Vec3 0 0 0, p1 = Vec3 1 0 0, p2 = Vec3 0 1 0, n0 = Vec3 0 0 1, '
                               'n1 = Vec3 0 0 1, n2 = Vec3 0 0 (-1), materialId = "fooMaterial" }') + ")")
    assert sc["sph_center"].tolist() == [[1, -2.5, F(0.3)]]
    assert sc["sph_radius"].tolist() == [2]
    assert sc["tri_n2"].tolist() == [[0, 0, -1]]
    assert (sc["prim_order"] >> 31).tolist() == [0, 1]


def _tri(k):
    return ('Triangle { p0 = Vec3 %d 0 0, p1 = Vec3 1 0 0, p2 = Vec3 0 1 0, n0 = Vec3 0 0 1, n1 = Vec3 0 0 1, '
            'n2 = Vec3 0 0 1, materialId = "fooMaterial" }' % k)


@pytest.mark.parametrize("size,n", [(0.5, 230), (0.25, 230), (1.0, 150), (0.004, 350), (0.0, 120)])
def test_subsampled_pick_semantics(size, n):
    """Scene.hs:80-85: take ceiling(size*100) of every 100 expanded objects."""
    sc, _, _ = api.load_yahrr(wrap('Subsampled { subsampleSize = %r, subsampledObjects = [ %s ] }'
                                   % (size, ", ".join(_tri(k) for k in range(n)))))
    of_hundred = int(np.ceil(F(size) * F(100)))
    want = [k for k in range(n) if (k % 100) < of_hundred] if of_hundred < 100 else list(range(n))
    assert sc["tri_p0"][:, 0].astype(int).tolist() == want


@pytest.mark.parametrize("text,code", [
    ("Scene { }", 9),
    (wrap('Sphere { position = Vec3 0 0 0, radius = 1 }'), 9),                       # missing field
    (wrap('Sphere { radius = 1, position = Vec3 0 0 0, materialId = "fooMaterial" }'), 9),   # field order matters
    (wrap('Sphere { position = Vec3 0 0 0, radius = 1, materialId = "fooMaterial" }') + " trailing", 9),
    (wrap('Sphere { position = Vec3 0 0 0, radius = 1, materialId = "nope" }'), 5),  # Map.! failure (main.hs:55)
    (wrap('TriangleMesh { triangleMeshPoints = [Vec3 0 0 0], triangleMeshNormals = Nothing, '
          'triangleMeshTriangles = [(0, 1, 2)], triangleMeshSmooth = Nothing, materialId = "fooMaterial" }'), 9),
])
def test_rejected_inputs(text, code):
    with pytest.raises(api.YahrError) as e:
        api.load_yahrr(text)
    assert e.value.code == code


def test_quantisation_matches_juicypixels_formula():
    x = np.array([-1.0, -0.0, 0.0, 0.001, 0.5, 0.999, 1.0, 1.5, np.inf, -np.inf, np.nan, 0.00392156, 0.0039216], F)
    q = api.quantize_rgb8_host(x)
    with np.errstate(invalid="ignore"):
        want = np.where(np.isnan(x), 0, np.trunc(F(255) * np.clip(x, 0, 1))).astype(np.uint8)
    assert q.tolist() == want.tolist()


def test_png_writer_roundtrip(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    for (h, w) in [(1, 1), (7, 13), (300, 257)]:       # the last one spans several stored deflate blocks
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        p = str(tmp_path / ("t%dx%d.png" % (w, h)))
        api.write_png_rgb8(p, img)
        back = np.asarray(Image.open(p).convert("RGB"))
        assert np.array_equal(back, img)
