"""Pins the CPU oracle against every known-answer / property test of the reference's own
test/Spec.hs (the only golden material the reference holds for this path).
Each test cites the Spec.hs lines it restates.  Tolerance: `near a b = len (a - b) < 1e-4`
(Spec.hs:23-24)."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import binding as ob

F = np.float32
_f32p = C.POINTER(C.c_float)


def P(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return a, a.ctypes.data_as(_f32p)


def near(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return float(np.sqrt(((a - b) ** 2).sum())) < 1e-4


def passes_through(x, x0, u):  # Spec.hs:26-31
    x, x0, u = (np.asarray(v, np.float64) for v in (x, x0, u))
    au = np.abs(u)
    d = 0 if (au[0] > au[1] and au[0] > au[2]) else (1 if au[1] > au[2] else 2)
    t = (x[d] - x0[d]) / u[d]
    return np.linalg.norm(x - x0 - t * u) < 1e-4


def cross(a, b):
    L = ob.lib()
    _, pa = P(a)
    _, pb = P(b)
    out = np.zeros(3, F)
    L.yo_cross(pa, pb, out.ctypes.data_as(_f32p))
    return out


def normv(a):
    L = ob.lib()
    _, pa = P(a)
    out = np.zeros(3, F)
    L.yo_norm(pa, out.ctypes.data_as(_f32p))
    return out


# ---- Vector (Spec.hs:35-47) ------------------------------------------------------------
def test_cross_forward_x_up_is_left():
    assert near(cross([0, 0, 1], [0, 1, 0]), [-1, 0, 0])


def test_cross_up_x_forward_is_right():
    assert near(cross([0, 1, 0], [0, 0, 1]), [1, 0, 0])


def test_cross_left_x_forward_is_up():
    assert near(cross([-1, 0, 0], [0, 0, 1]), [0, 1, 0])


# ---- Ray / collideAll (Spec.hs:49-60) --------------------------------------------------
def test_collide_all_last_just_wins():
    L = ob.lib()
    present = (C.c_int * 3)(0, 1, 1)
    ts = (C.c_float * 3)(0, 2, 1)
    whats = (C.c_int * 3)(0, 2, 1)
    assert L.yo_collide_all_const(3, present, ts, whats) == 1
    # all Nothing -> Nothing
    assert L.yo_collide_all_const(3, (C.c_int * 3)(0, 0, 0), ts, whats) == -1


# ---- Camera (Spec.hs:62-108) -----------------------------------------------------------
ORIGIN = [10, -80, 124]


def cam_ray(up, look, u, v):
    L = ob.lib()
    cam = ob.make_camera(dict(imW=200, imH=200, focalLength=1, lookDir=look, upDir=up, position=ORIGIN))
    out = np.zeros(7, F)
    L.yo_camera_ray(C.byref(cam), u, v, out.ctypes.data_as(_f32p))
    return out[:3], out[3:6], out[6]


@pytest.mark.parametrize("up,look,corner", [
    ([0, 1, 0], [0, 0, 1], [-0.5, 0.5, 1]),      # looking forward  (Spec.hs:70-82)
    ([0, 0, 1], [0, -1, 0], [-0.5, -1, 0.5]),    # looking down     (Spec.hs:84-95)
    ([0, 1, 0], [1, 0, 0], [1, 0.5, 0.5]),       # looking right    (Spec.hs:97-108)
])
def test_camera(up, look, corner):
    x0, u, tmax = cam_ray(up, look, 100, 100)
    assert near(x0, ORIGIN)
    assert near(u, look)
    assert passes_through(np.add(ORIGIN, look), x0, u)
    assert tmax == F(1e6)
    x0, u, _ = cam_ray(up, look, 0, 0)
    assert near(x0, ORIGIN)
    assert passes_through(np.add(ORIGIN, corner), x0, u)


def test_camera_origin_is_position_exactly():
    # tf . (0,0,0,1) reproduces `position` bit-exactly
    x0, _, _ = cam_ray([0, 1, 0], [0.4, -0.3, 1], 17, 5)
    assert np.array_equal(x0, np.asarray(ORIGIN, F))


# ---- Shapes (Spec.hs:110-145) ----------------------------------------------------------
def test_sphere_bound():
    L = ob.lib()
    out = np.zeros(6, F)
    _, ps = P([10, 10, 15])
    L.yo_bound_sphere(5.0, ps, out.ctypes.data_as(_f32p))
    assert near(out[:3], [5, 5, 10]) and near(out[3:], [15, 15, 20])


TRI = ([0, 0, 5], [2, 0, 5], [1, 2, 5])
TRI_N = [0, 0, -1]


def tri_hit(x, y, n=TRI_N, tmax=1e6):
    L = ob.lib()
    ptrs = [P(v)[1] for v in TRI] + [P(n)[1]] * 3
    keep = [P(v)[0] for v in TRI]
    out = np.zeros(13, F)
    a, px = P([x, y, 0])
    b, pu = P([0, 0, 1])
    ok = L.yo_collide_triangle(*ptrs, px, pu, tmax, out.ctypes.data_as(_f32p))
    return ok, out


@pytest.mark.parametrize("x,y", [(1, 1), (0.1, 0.1), (0.9, 0.1)])
def test_triangle_hits(x, y):
    ok, out = tri_hit(x, y)
    assert ok == 1
    assert near(out[1:4], [x, y, 5])
    assert near(out[4:7], [0, 0, -1])


@pytest.mark.parametrize("x,y", [(0, 1), (-0.1, 1), (2, 1)])
def test_triangle_misses(x, y):
    ok, _ = tri_hit(x, y)
    assert ok == 0


def test_triangle_shading_normal_backface_is_a_miss():
    # (u . ns) < 0 is part of the intersection test (Shapes.hs:57)
    ok, _ = tri_hit(1, 1, n=[0, 0, 1])
    assert ok == 0


def test_triangle_tmax_is_inclusive():
    ok, out = tri_hit(1, 1, tmax=5.0)       # t <= tMax (Shapes.hs:57)
    assert ok == 1 and out[0] == F(5.0)
    ok, _ = tri_hit(1, 1, tmax=float(np.nextafter(F(5.0), F(0))))
    assert ok == 0


def test_triangle_bound():
    L = ob.lib()
    out = np.zeros(6, F)
    L.yo_bound_triangle(*[P(v)[1] for v in TRI], out.ctypes.data_as(_f32p))
    assert near(out[:3], [0, 0, 5]) and near(out[3:], [2, 2, 5])


# ---- AABB (Spec.hs:176-218) ------------------------------------------------------------
def test_bb_from_points():
    L = ob.lib()
    out = np.zeros(6, F)
    L.yo_bb_from_points(P([1, 20, 5])[1], P([14, 4, 3])[1], out.ctypes.data_as(_f32p))
    assert near(out[:3], [1, 4, 3]) and near(out[3:], [14, 20, 5])


def test_bb_include_point():
    L = ob.lib()
    out = np.zeros(6, F)
    L.yo_bb_include_point(P([1, 4, 3, 14, 20, 5])[1], P([0, 22, 3])[1], out.ctypes.data_as(_f32p))
    assert near(out[:3], [0, 4, 3]) and near(out[3:], [14, 22, 5])


def test_bb_join():
    L = ob.lib()
    out = np.zeros(6, F)
    L.yo_bb_join(P([1, 4, 3, 14, 20, 5])[1], P([0, 4, 2, 15, 20, 1])[1], out.ctypes.data_as(_f32p))
    assert near(out[:3], [0, 4, 2]) and near(out[3:], [15, 20, 5])


BB = [10, 20, 30, 20, 30, 40]


def bb_ray(x0, d):
    L = ob.lib()
    tn = C.c_float(0)
    u = normv(d)
    return L.yo_bb_ray(P(BB)[1], P(x0)[1], P(u)[1], 1e6, C.byref(tn)), tn.value


@pytest.mark.parametrize("x0,d", [([9, 29, 35], [1, 1.1, 0]), ([21, 29, 35], [-1, 1.1, 0]), ([15, 19, 39], [0, 1, 1.1])])
def test_bb_missed(x0, d):
    assert bb_ray(x0, d)[0] == 0


@pytest.mark.parametrize("x0,d", [([9, 29, 35], [1, 0.9, 0]), ([21, 29, 35], [-1, 0.9, 0]), ([15, 19, 39], [0, 1, 0.9])])
def test_bb_hit(x0, d):
    assert bb_ray(x0, d)[0] == 1


def test_bb_centroid():
    L = ob.lib()
    out = np.zeros(3, F)
    L.yo_bb_centroid(P([-10, -10, -10, 10, 10, 10])[1], out.ctypes.data_as(_f32p))
    assert near(out, [0, 0, 0])


def test_bb_nan_passes():
    # Ray origin exactly on a slab plane with a zero direction component: 0 * inf = NaN and the
    # reference's min/max selects end in "box passes" (AABBs.hs:40, SURVEY.md note N).
    # Hand-traced through AABBs.hs:29-40 with GHC's min/max selects:
    L = ob.lib()
    tn = C.c_float(0)
    # x0.x on bMin.x, u.x = +0: t0 = 0*inf = NaN, t1 = +inf -> tDimNear = t1 = +inf -> miss
    assert L.yo_bb_ray(P(BB)[1], P([10, 25, 0])[1], P([0.0, 0, 1])[1], 1e6, C.byref(tn)) == 0
    # x0.x on bMin.x, u.x = -0: t0 = NaN, t1 = -inf -> near = -inf (ignored), far = NaN poisons
    # tFar, the Y slab resets it to +inf, Z gives [30, 40] -> pass with tNear = 30
    assert L.yo_bb_ray(P(BB)[1], P([10, 25, 0])[1], P([-0.0, 0, 1])[1], 1e6, C.byref(tn)) == 1
    assert tn.value == 30.0
    # the poisoned tFar forgets tMax: the same ray with tMax = 1 (< 30) still passes
    assert L.yo_bb_ray(P(BB)[1], P([10, 25, 0])[1], P([-0.0, 0, 1])[1], 1.0, C.byref(tn)) == 1
    # x0.x on bMax.x, u.x = +0: t0 = -inf, t1 = NaN -> far = -inf -> miss
    assert L.yo_bb_ray(P(BB)[1], P([20, 25, 0])[1], P([0.0, 0, 1])[1], 1e6, C.byref(tn)) == 0
    # strictly inside the X slab with u.x = 0: t0 = -inf, t1 = +inf -> no constraint -> pass
    assert L.yo_bb_ray(P(BB)[1], P([15, 25, 0])[1], P([0.0, 0, 1])[1], 1e6, C.byref(tn)) == 1


# ---- Culling / BVH (Spec.hs:221-262) ---------------------------------------------------
def grid_scene(split_mode, depth=16):
    grid = [(i, j, k) for i in range(1, 11) for j in range(1, 11) for k in (1, 2)]
    centers = np.array([[i * 10, j * 10, k * 10] for (i, j, k) in grid], F)
    sc = dict(sph_center=centers, sph_radius=np.ones(len(grid), F), sph_material=np.zeros(len(grid), np.uint32),
              materials=np.array([[1, 1, 1, 0, 0, 0, 1]], F), lights=np.zeros((0, 6), F), bvh_max_depth=depth,
              split_mode=split_mode)
    return grid, sc


@pytest.mark.parametrize("split_mode", [1, 0])  # SAH (Spec.hs:226-243), Midpoint (Spec.hs:245-262)
def test_bvh_grid(split_mode):
    grid, sc = grid_scene(split_mode)
    s = ob.OracleScene(sc)
    for x, y, want in [(10, 10, (1, 1, 1)), (30, 30, (3, 3, 1)), (10, 100, (1, 10, 1)), (100, 10, (10, 1, 1)),
                       (50, 50, (5, 5, 1)), (10.49, 10.49, (1, 1, 1))]:
        pid, _ = s.collide([x, y, 0], [0, 0, 1])
        assert pid >= 0
        assert grid[pid] == want
    # every primitive appears exactly once in the DFS leaf order
    assert sorted(s.bvh_order().tolist()) == list(range(len(grid)))
    s.close()


@pytest.mark.parametrize("split_mode", [0, 1])
def test_bvh_equals_brute_force(split_mode):
    # a depth-0 BVH is one multi-leaf = linear collideAll over all primitives (Culling.hs:34)
    grid, sc = grid_scene(split_mode)
    _, sc0 = grid_scene(split_mode, depth=0)
    a, b = ob.OracleScene(sc), ob.OracleScene(sc0)
    rng = np.random.default_rng(1)
    for _ in range(300):
        x0 = rng.uniform(-5, 115, 3).astype(F)
        x0[2] = -5
        d = normv(rng.normal(size=3).astype(F) * F(0.2) + np.array([0, 0, 1], F))
        pa, oa = a.collide(x0, d)
        pb, ob_ = b.collide(x0, d)
        assert pa == pb
        if pa >= 0:
            assert oa[0] == ob_[0]
    a.close()
    b.close()


# ---- Integrator (Spec.hs:265-297) ------------------------------------------------------
def integ_scene(light_pos):
    return dict(
        tri_p0=[[-10, -10, 0]], tri_p1=[[10, -10, 0]], tri_p2=[[0, 10, 0]],
        tri_n0=[[0, 0, 1]], tri_n1=[[0, 0, 1]], tri_n2=[[0, 0, 1]], tri_material=[0],
        materials=[[0, 0, 0, 1, 1, 1, 10]],           # Blinn 10 only
        lights=[list(light_pos) + [1, 1, 1]], bvh_max_depth=16, split_mode=0)


def vec(angle):
    # computed in Float like the Spec (cos/sin of a Float angle)
    a = F(angle)
    return np.array([np.cos(a, dtype=F), 0, np.sin(a, dtype=F)], F)


def pt(angle):
    return -vec(angle)


PI = F(math.pi)


def test_integrator_rear_side_is_black():
    s = ob.OracleScene(integ_scene(vec(PI / F(4))))
    assert near(s.radiance([0, 0, -1], [0, 0, 1]), [0, 0, 0])
    s.close()
    s = ob.OracleScene(integ_scene(pt(F(0.25) * PI)))
    a = F(1.5) * PI
    assert near(s.radiance(pt(a), vec(a)), [0, 0, 0])
    s.close()


def test_integrator_mirror_direction_monotone():
    a = F(0.25) * PI
    rads = []
    for la in (0.25, 0.50, 0.75):
        s = ob.OracleScene(integ_scene(pt(F(la) * PI)))
        L = s.radiance(pt(a), vec(a))
        rads.append(float((L.astype(np.float64) ** 2).sum()))
        s.close()
    assert rads == sorted(rads)
    # NB: literally as in Spec.hs the eye ray arrives from z < 0 with u.n > 0, so the shading-normal
    # test (Shapes.hs:57) rejects the hit and all three radiances are 0: the reference's check is
    # vacuous.  The non-vacuous variant (eye above the surface) follows.
    assert rads == [0.0, 0.0, 0.0]


def test_integrator_mirror_direction_monotone_from_above():
    a = F(0.25) * PI
    eye_x0 = np.array([-np.cos(a, dtype=F), 0, np.sin(a, dtype=F)], F)
    eye_u = np.array([np.cos(a, dtype=F), 0, -np.sin(a, dtype=F)], F)
    rads = []
    for la in (0.75, 0.50, 0.25):  # light moves towards the mirror direction (angle 0.25 pi above +x)
        lp = np.array([np.cos(F(la) * PI, dtype=F), 0, np.sin(F(la) * PI, dtype=F)], F)
        s = ob.OracleScene(integ_scene(lp))
        L = s.radiance(eye_x0, eye_u)
        rads.append(float((L.astype(np.float64) ** 2).sum()))
        s.close()
    assert rads == sorted(rads)
    assert rads[0] > 0


# ---- BSDF (Spec.hs:300-328) ------------------------------------------------------------
DG = [0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 1, 0]


def blinn(light, eye, e=10.0):
    L = ob.lib()
    out = np.zeros(3, F)
    L.yo_bsdf_at(1, P([e])[1], P(DG)[1], P(light)[1], P(eye)[1], out.ctypes.data_as(_f32p))
    return out


def test_blinn_rear_side():
    assert near(blinn(vec(F(5) * PI / F(4)), vec(F(7) * PI / F(4))), [0, 0, 0])


def test_blinn_no_transmission():
    eye = vec(F(0.25) * PI)
    for la in (1.25, 1.5, 1.75):
        assert near(blinn(vec(F(la) * PI), eye), [0, 0, 0])


def test_blinn_monotone_and_values():
    eye = vec(F(0.25) * PI)
    fs = [float((blinn(vec(F(la) * PI), eye).astype(np.float64) ** 2).sum()) for la in (0.25, 0.5, 0.75, 0.9)]
    assert fs == sorted(fs)
    # float32 restatement gives 0.02984, 0.30592, 0.95493, 1.02047 (per channel)
    vals = [float(blinn(vec(F(la) * PI), eye)[0]) for la in (0.25, 0.5, 0.75, 0.9)]
    assert np.allclose(vals, [0.02984, 0.30592, 0.95493, 1.02047], atol=2e-4)


def test_lambertian_is_one_over_pi():
    L = ob.lib()
    out = np.zeros(3, F)
    L.yo_bsdf_at(0, P([0])[1], P(DG)[1], P(vec(F(0.3) * PI))[1], P(vec(F(0.6) * PI))[1], out.ctypes.data_as(_f32p))
    assert np.array_equal(out, np.full(3, F(1) / F(np.float32(3.14159274101257324)), F))


# ---- Sampling (Spec.hs:330-346) --------------------------------------------------------
@pytest.mark.parametrize("w,h,n", [(40, 40, 5), (40, 40, 16), (1, 1, 1), (20, 1, 3), (1, 20, 3)])
def test_square_batches_cover_without_overlap(w, h, n):
    L = ob.lib()
    seen = np.zeros((h, w), np.int32)
    win = (C.c_int32 * 4)()
    for i in range(n):
        L.yo_batch_window(w, h, i, n, win)
        x0, y0, x1, y1 = win
        seen[y0:y1, x0:x1] += 1
    assert (seen == 1).all()


def test_num_batches_matches_survey_table():
    # 512^2 -> 1024 tiles; 1920x1080 -> 8192; 2048^2 -> 16384; 3840x2160 -> 32768
    L = ob.lib()
    assert L.yo_num_batches(8, 512, 512) == 1024
    assert L.yo_num_batches(8, 1920, 1080) == 8192
    assert L.yo_num_batches(8, 2048, 2048) == 16384
    assert L.yo_num_batches(8, 3840, 2160) == 32768


# ---- Scene.expand (Spec.hs:148-172) is host-side; see tests/test_host_scene.py ----------
