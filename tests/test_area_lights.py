"""Quad area lights -- an EXTENSION with no reference counterpart (the reference has point lights only,
Lights.hs:7; SURVEY.md F3: "parity unpinned").  The oracle defines the semantics (oracle/yahr_oracle.hpp,
`lightSampleU`, `finishAreaLight`, `directIllumination`); these CPU tests pin the oracle's behaviour through
properties, the GPU tests (test_gpu_parity.py) hold the kernels to the oracle."""
import copy

import numpy as np

from oracle import binding as ob
from yahr_b200 import scenes

F = np.float32


def floor_scene():
    sc = scenes._empty_scene()
    fp0, fp1, fp2, fn = scenes._floor(0.0, half=5.0)
    for k, a in zip(("tri_p0", "tri_p1", "tri_p2", "tri_n0", "tri_n1", "tri_n2"), (fp0, fp1, fp2, fn, fn, fn)):
        sc[k] = a
    sc["tri_material"] = np.zeros(2, np.uint32)
    sc["materials"] = np.array([[0.8, 0.8, 0.8, 0, 0, 0, 1]], F)
    cam = scenes._camera(64, 64, 1.0, [0, -1, 0.2], [0, 0, 1], [0, 6, -1])
    return sc, cam


def render(sc, cam, **kw):
    o = ob.OracleScene(sc)
    rgb, pid, _, st = o.render(cam, **kw)
    o.close()
    return rgb, pid, st


def test_small_distant_area_light_is_a_point_light_times_the_emitter_cosine():
    sc, cam = floor_scene()
    sa = scenes.add_area_light(copy.deepcopy(sc), [-0.005, 4.0, -0.005], [0.01, 0, 0], [0, 0, 0.01], [1e5, 1e5, 1e5], 8)
    sp = copy.deepcopy(sc)
    sp["lights"] = np.array([[0, 4.0, 0, 10, 10, 10]], F)        # radiance * area = 1e5 * 1e-4
    ra, pa, sta = render(sa, cam)
    rp, pp, stp = render(sp, cam)
    assert np.array_equal(pa, pp)
    assert sta["n_shadow"] == 8 * stp["n_shadow"]
    # directly below the panel the emitter cosine is 1
    assert abs(ra[32, 32, 0] - rp[32, 32, 0]) <= 0.005 * rp[32, 32, 0]
    # elsewhere the area light is dimmer by cos_l <= 1, never brighter
    hit = pa != 0xFFFFFFFF
    assert (ra[hit] <= rp[hit] * 1.001 + 1e-7).all()
    assert ra[hit].mean() < rp[hit].mean()


def test_area_light_is_one_sided():
    sc, cam = floor_scene()
    down = scenes.add_area_light(copy.deepcopy(sc), [-1, 4, -1], [2, 0, 0], [0, 0, 2], [5, 5, 5], 4)
    up = scenes.add_area_light(copy.deepcopy(sc), [-1, 4, -1], [0, 0, 2], [2, 0, 0], [5, 5, 5], 4)   # edges swapped
    rd, _, std = render(down, cam)
    ru, _, stu = render(up, cam)
    assert rd.max() > 0 and std["n_shadow"] > 0
    assert (ru == 0).all() and stu["n_shadow"] == 0


def test_area_light_sampling_is_counter_based():
    sc, cam = floor_scene()
    big = scenes.add_area_light(copy.deepcopy(sc), [-2, 3, -2], [4, 0, 0], [0, 0, 4], [2, 2, 2], 2)
    a, _, _ = render(big, cam, seed=1)
    b, _, _ = render(big, cam, seed=1)
    c, _, _ = render(big, cam, seed=2)
    assert np.array_equal(a, b)                       # same seed: identical
    assert not np.array_equal(a, c)                   # another seed: other sample points
    # more samples: same mean brightness (the flux is divided by the sample count), less noise
    many = scenes.add_area_light(copy.deepcopy(sc), [-2, 3, -2], [4, 0, 0], [0, 0, 4], [2, 2, 2], 32)
    m, pm, _ = render(many, cam, seed=1)
    hit = pm != 0xFFFFFFFF
    assert abs(m[hit].mean() - a[hit].mean()) < 0.05 * m[hit].mean()


def test_area_light_is_occluded_by_geometry():
    sc, cam = floor_scene()
    lit = scenes.add_area_light(copy.deepcopy(sc), [-0.5, 4, -0.5], [1, 0, 0], [0, 0, 1], [20, 20, 20], 4)
    blocked = copy.deepcopy(lit)
    blocked["sph_center"] = np.array([[0, 2.0, 0]], F)     # a sphere between the panel and the floor centre
    blocked["sph_radius"] = np.array([1.0], F)
    blocked["sph_material"] = np.zeros(1, np.uint32)
    a, pa, _ = render(lit, cam)
    b, pb, _ = render(blocked, cam)
    # primitive IDs: spheres first, so the sphere is 0 in `blocked` and the floor triangles move to 1, 2
    floor_both = (pa != 0xFFFFFFFF) & (pb != 0xFFFFFFFF) & (pb != 0)
    assert (b[floor_both] <= a[floor_both]).all()
    assert b[floor_both].sum() < 0.9 * a[floor_both].sum()
