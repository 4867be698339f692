"""GPU BVH construction against the host builder, which itself is checked
bit-for-bit against the oracle's tree in tests/test_host_logic.py: same left-first DFS primitive order
(= tie-break ranks), same pre-order node numbering, same boxes, same child references."""
import os

import numpy as np
import pytest

from yahr_b200 import api, scenes

pytestmark = pytest.mark.gpu


def build(sc, host):
    if host:
        os.environ["YAHR_B200_HOST_BUILD"] = "1"
    else:
        os.environ.pop("YAHR_B200_HOST_BUILD", None)
    try:
        s = api.Scene(sc)
    finally:
        os.environ.pop("YAHR_B200_HOST_BUILD", None)
    out = (s.info(), s.download_bvh())
    s.close()
    return out


def resolve_refs(nodes, multi):
    """Child references with multi-leaf table indices replaced by their (first, count) payload."""
    refs = nodes[:, 12:14].copy().view(np.uint32)
    out = []
    for r in refs.reshape(-1):
        if (r & 0xC0000000) == 0xC0000000 and r != 0xFFFFFFFF:
            f, c = multi[r & 0x3FFFFFFF]
            out.append(("multi", int(f), int(c)))
        else:
            out.append(("ref", int(r), 0))
    return out


CASES = {
    "c1": lambda: scenes.c1_scene_yahrr()[0],
    "bunny": lambda: scenes.c2_bunny_proxy(64, 36)[0],                      # 69 566 triangles
    "bunny-depth9": lambda: dict(scenes.c2_bunny_proxy(64, 36, nu=60, nv=40)[0], bvh_max_depth=9),
    "grid": lambda: scenes.c3_sphere_grid(16, 64, 64)[0],
    "grid-depth0": lambda: dict(scenes.c3_sphere_grid(6, 64, 64)[0], bvh_max_depth=0),
    "terrain": lambda: scenes.c4_terrain(301, 151, 64, 36)[0],
    "soup": lambda: scenes.c4_soup(100_000, 64, 36)[0],
    "adversarial": lambda: scenes.adversarial_shared_edges()[0],             # coplanar duplicates, mixed kinds
    "tiny-1": lambda: dict(scenes.c1_scene_yahrr()[0], sph_center=scenes.c1_scene_yahrr()[0]["sph_center"][:1],
                           sph_radius=np.ones(1, np.float32), sph_material=np.zeros(1, np.uint32),
                           tri_p0=np.zeros((0, 3)), tri_p1=np.zeros((0, 3)), tri_p2=np.zeros((0, 3)),
                           tri_n0=np.zeros((0, 3)), tri_n1=np.zeros((0, 3)), tri_n2=np.zeros((0, 3)),
                           tri_material=np.zeros(0, np.uint32)),
}


@pytest.mark.parametrize("split_mode", [0, 1])         # Midpoint, SurfaceAreaHeuristic (Culling.hs:18)
@pytest.mark.parametrize("name", list(CASES.keys()))
def test_device_build_equals_host_build(name, split_mode):
    sc = dict(CASES[name](), split_mode=split_mode)
    hi, (horder, hnodes, hmulti, hroot, hbox) = build(sc, host=True)
    di, (dorder, dnodes, dmulti, droot, dbox) = build(sc, host=False)
    assert hi["built_on_device"] == 0 and di["built_on_device"] == 1
    assert np.array_equal(dorder, horder), "DFS primitive order differs"
    for k in ("n_primitives", "n_nodes", "n_multi_leaves", "depth"):
        assert di[k] == hi[k], k
    assert np.array_equal(dnodes[:, :12], hnodes[:, :12]), "child boxes differ"     # float ==: -0 equals +0
    assert resolve_refs(dnodes, dmulti) == resolve_refs(hnodes, hmulti)
    assert np.array_equal(dbox, hbox)
    if (hroot & 0xC0000000) == 0xC0000000 and hroot != 0xFFFFFFFF:
        assert tuple(dmulti[droot & 0x3FFFFFFF]) == tuple(hmulti[hroot & 0x3FFFFFFF])
    else:
        assert droot == hroot


def test_device_build_full_size_c4_and_timing():
    sc = scenes.c4_terrain()[0]
    hi, (horder, hnodes, hmulti, _, _) = build(sc, host=True)
    di, (dorder, dnodes, dmulti, _, _) = build(sc, host=False)
    assert np.array_equal(dorder, horder)
    assert np.array_equal(dnodes[:, :12], hnodes[:, :12])
    assert resolve_refs(dnodes, dmulti) == resolve_refs(hnodes, hmulti)
    print("BVH build 1M triangles: host %.1f ms, device %.1f ms (+ %.1f ms upload)" %
          (hi["build_ms"], di["build_ms"], di["upload_ms"]))
    assert di["build_ms"] < hi["build_ms"]


def test_device_build_errors():
    sc = dict(scenes.c1_scene_yahrr()[0])
    sc["sph_material"] = np.full(7, 99, np.uint32)
    with pytest.raises(api.YahrError) as e:
        api.Scene(sc)
    assert e.value.code == 5
    sc = dict(scenes.c1_scene_yahrr()[0])
    c = sc["sph_center"].copy()
    c[3, 1] = np.inf
    sc["sph_center"] = c
    with pytest.raises(api.YahrError) as e:
        api.Scene(sc)
    assert e.value.code == 7


def test_sah_scenes_are_built_on_the_device_too():
    sc = dict(scenes.c3_sphere_grid(8, 64, 64)[0], split_mode=1)
    s = api.Scene(sc)
    assert s.info()["built_on_device"] == 1
    s.close()


def test_device_sah_build_full_size_c4_soup():
    """1 M random triangles, SurfaceAreaHeuristic: the bucket counts / boxes come from atomics, the cost expression
    from one thread per segment -- the tree must still be the host builder's (= the oracle's) bit for bit."""
    sc = dict(scenes.c4_soup()[0], split_mode=1)
    hi, (horder, hnodes, hmulti, _, _) = build(sc, host=True)
    di, (dorder, dnodes, dmulti, _, _) = build(sc, host=False)
    assert di["built_on_device"] == 1
    assert np.array_equal(dorder, horder)
    assert np.array_equal(dnodes[:, :12], hnodes[:, :12])
    assert resolve_refs(dnodes, dmulti) == resolve_refs(hnodes, hmulti)
    print("SAH BVH build 1M triangles: host %.1f ms, device %.1f ms" % (hi["build_ms"], di["build_ms"]))
