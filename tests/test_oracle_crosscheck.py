"""Second, independent restatement of the reference's per-ray functions, used to cross-check the C++ oracle.

The C++ oracle (oracle/yahr_oracle.hpp) is what the GPU path is held to; the real reference cannot be run here
(no GHC).  To keep a misreading of the Haskell from going unnoticed, the functions below restate the same Haskell
definitions a second time, straight from the source files, in numpy binary32 scalar arithmetic (every operation
rounds to float32, no contraction), with the operator precedences of Vectors.hs:34,44 (`infixl 7 .*`, `infixl 7 @*`)
and GHC's class-default min / max.  The two restatements must agree BIT FOR BIT on hit / miss, t and the
differential geometry for random and adversarial inputs (the Blinn lobe, which goes through powf, within 2 ulp).

  bbRayIntersection  AABBs.hs:29-40        collideSphere    Shapes.hs:13-29
  collideTriangle    Shapes.hs:36-59       BSDF.at          BSDF.hs:12-46
  computeInitialRay  Cameras.hs:59-86 (+ linear's !*!, !*: row sums folded from 0, left to right)
"""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob

F = np.float32


def f(x):
    return F(x)


# ---- Vectors.hs ------------------------------------------------------------------------------------------
def vsub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def vadd(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def vmul(a, b):                       # Num Vec3: component-wise
    return (a[0] * b[0], a[1] * b[1], a[2] * b[2])


def dot(a, b):                        # (.*) = x1*y1 + x2*y2 + x3*y3, left to right          Vectors.hs:32-33
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def scale(s, v):                      # (@*)                                                   Vectors.hs:42-43
    return (s * v[0], s * v[1], s * v[2])


def cross(a, b):                      # Vectors.hs:49-53
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def norm(v):                          # norm v = (1 / len v) @* v ; len = sqrt . lensq         Vectors.hs:46-47,56-60
    return scale(F(1) / np.sqrt(dot(v, v)), v)


def hmin(x, y):                       # Ord Float class defaults
    return x if x <= y else y


def hmax(x, y):
    return y if x <= y else x


def V(x):
    return tuple(F(t) for t in x)


# ---- AABBs.hs:29-40 --------------------------------------------------------------------------------------
def bb_ray(bmin, bmax, x0, u, tmax):
    tnear, tfar = F(0), F(tmax)
    for d in range(3):
        inv = F(1) / u[d]
        t0 = (bmin[d] - x0[d]) * inv
        t1 = (bmax[d] - x0[d]) * inv
        tnear, tfar = hmax(tnear, hmin(t0, t1)), hmin(tfar, hmax(t0, t1))
    return (not (tnear > tfar)), tnear


# ---- Shapes.hs:13-29 -------------------------------------------------------------------------------------
def collide_sphere(r, s, x0, u, tmax):
    d = vsub(x0, s)
    a = dot(u, u)
    b = dot(vmul(V((2, 2, 2)), d), u)          # 2 * d .* u  =  (2 * d) .* u   (both infixl 7; 2 :: Vec3 = vof 2)
    c = dot(d, d) - r * r
    delta = b * b - (F(4) * a) * c
    if delta < 0:
        return None
    sq = np.sqrt(delta)
    for t in ((-b - sq) / (F(2) * a), (-b + sq) / (F(2) * a)):
        if t > 0 and t <= tmax:
            x = vadd(x0, scale(t, u))
            n = norm(vsub(x, s))
            return t, x, n, cross(n, V((0, 0, 1))), cross(n, V((0, 1, 0)))
    return None


# ---- Shapes.hs:36-59 -------------------------------------------------------------------------------------
def collide_triangle(p0, p1, p2, n0, n1, n2, x0, u, tmax):
    e1, e2, s = vsub(p1, p0), vsub(p2, p0), vsub(x0, p0)
    s1, s2 = cross(u, e2), cross(s, e1)
    inv = F(1) / dot(s1, e1)
    t = dot(s2, e2) * inv
    b1 = dot(s1, s) * inv
    b2 = dot(s2, u) * inv
    b0 = (F(1) - b1) - b2
    ns = vadd(vadd(scale(b0, n0), scale(b1, n1)), scale(b2, n2))
    ok = (b0 >= 0 and b0 <= 1 and b1 >= 0 and b1 <= 1 and b2 >= 0 and b2 <= 1 and t > 0 and t <= tmax and dot(u, ns) < 0)
    if not ok:
        return None
    ss = norm(e2)
    ts = norm(cross(ss, ns))
    return t, vadd(x0, scale(t, u)), ns, cross(ts, ns), ts


# ---- BSDF.hs:12-46 with Shaders.blinnPhong (Shaders.hs:12-14) ---------------------------------------------
def bsdf_at(diffuse, specular, exponent, normal, dpdu, i_w, o_w):
    nn = normal
    sn = norm(dpdu)
    tn = cross(nn, sn)

    def loc(v):
        return (dot(v, sn), dot(v, tn), dot(v, nn))

    i, o = loc(i_w), loc(o_w)
    pi = F(np.pi)
    if i[2] > 0 and o[2] > 0:
        lam = F(1) / pi
        h = norm(vadd(i, o))
        cos_o, cos_i = abs(o[2]), abs(i[2])
        cos_h, o_dot_h = dot(i, h), dot(o, h)
        d = ((exponent + F(2)) * F(np.float32(abs(h[2])) ** np.float32(exponent))) / (F(2) * pi)
        g = hmin(F(1), hmin(((F(2) * cos_h) * cos_o) / o_dot_h, ((F(2) * cos_h) * cos_i) / o_dot_h))
        bl = ((d * g) * F(1)) / ((F(4) * cos_i) * cos_o)
    else:
        lam, bl = F(0), F(0)
    a = vmul(diffuse, (lam, lam, lam))
    b = vmul(specular, (bl, bl, bl))
    return vadd(vadd(V((0, 0, 0)), a), b)        # sum = foldl (+) 0 over [Scaled diffuse Lambertian, Scaled specular Blinn]


# ---- Cameras.hs:59-86 ------------------------------------------------------------------------------------
def mat_mul(a, b):        # linear's !*! : row i = sum_k a[i][k] *^ b[k], folded from 0 left to right
    return [[((((F(0) + a[i][0] * b[0][j]) + a[i][1] * b[1][j]) + a[i][2] * b[2][j]) + a[i][3] * b[3][j]) for j in range(4)]
            for i in range(4)]


def mat_vec(m, v):        # linear's !* : row . v, same fold
    return [(((F(0) + m[i][0] * v[0]) + m[i][1] * v[1]) + m[i][2] * v[2]) + m[i][3] * v[3] for i in range(4)]


def camera_ray(imw, imh, focal, look, up, pos, u, v):
    forward = norm(look)
    left = norm(cross(forward, up))
    up2 = cross(left, forward)
    nl = (-left[0], -left[1], -left[2])
    o, z = F(1), F(0)
    dir_tf = [[nl[0], up2[0], forward[0], z], [nl[1], up2[1], forward[1], z], [nl[2], up2[2], forward[2], z], [z, z, z, o]]
    tr = [[o, z, z, pos[0]], [z, o, z, pos[1]], [z, z, o, pos[2]], [z, z, z, o]]
    tf = mat_mul(tr, dir_tf)
    fw, fh = imw / imh, F(1)
    r2c = [[fw / imw, z, z, -(fw / F(2))], [z, -(fh / imh), z, fh / F(2)], [z, z, o, z], [z, z, z, o]]
    vtf = mat_mul(tf, r2c)
    og = mat_vec(tf, [z, z, z, o])
    origin = scale(F(1) / og[3], (og[0], og[1], og[2]))          # normalizePoint
    p4 = mat_vec(vtf, [F(u), F(v), focal, o])
    p = scale(F(1) / p4[3], (p4[0], p4[1], p4[2]))
    return origin, norm(vsub(p, origin))


# ---- helpers ---------------------------------------------------------------------------------------------
def bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def same_bits(a, b):
    a, b = np.asarray(a, np.float32).ravel(), np.asarray(b, np.float32).ravel()
    nan = np.isnan(a) & np.isnan(b)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | nan | ((a == 0) & (b == 0))))


def arr(v):
    return np.ascontiguousarray(np.asarray(v, np.float32))


def ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.fixture(scope="module")
def L():
    ob.build()
    return ob.lib()


def random_dirs(rng, n):
    d = rng.normal(size=(n, 3)).astype(np.float32)
    # adversarial: exact zeros (1/0 = inf, 0 * inf = NaN in the slab test), axis-aligned, tiny components
    d[::7, 0] = 0
    d[::11, 1] = 0
    d[::13] = np.array([0, 0, 1], np.float32)
    d[::17, 2] = np.float32(1e-30)
    return d


def test_slab_test_agrees_bit_for_bit(L):
    rng = np.random.default_rng(1)
    n = 4000
    lo = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    hi = lo + rng.uniform(0, 4, (n, 3)).astype(np.float32)
    hi[::5, 1] = lo[::5, 1]                                   # flat boxes (the floor of scene.yahrr)
    x0 = rng.uniform(-8, 8, (n, 3)).astype(np.float32)
    x0[::9, 1] = lo[::9, 1]                                   # origin exactly on a slab plane
    u = random_dirs(rng, n)
    aim = (lo + (hi - lo) * rng.uniform(0, 1, (n, 3)).astype(np.float32)).astype(np.float32)
    toward = (aim - x0).astype(np.float32)
    u[1::3] = toward[1::3]                                    # a third of the rays aim at their box (unnormalised)
    u[2::3] = toward[2::3] / np.linalg.norm(toward[2::3], axis=1, keepdims=True).astype(np.float32)
    u[::21, 0] = 0                                            # ... some of them with an exactly zero component
    tmax = rng.choice(np.array([1e6, 3.0, 0.5], np.float32), n)
    hits = 0
    with np.errstate(all="ignore"):
        for k in range(n):
            ok, tn = bb_ray(V(lo[k]), V(hi[k]), V(x0[k]), V(u[k]), F(tmax[k]))
            out = np.zeros(1, np.float32)
            bb = arr(np.concatenate([lo[k], hi[k]]))
            got = L.yo_bb_ray(ptr(bb), ptr(arr(x0[k])), ptr(arr(u[k])), C.c_float(float(tmax[k])), ptr(out))
            assert bool(got) == bool(ok), k
            if ok:
                hits += 1
                assert same_bits(out[0], tn), k
    assert 200 < hits < n - 200


def test_triangle_test_agrees_bit_for_bit(L):
    rng = np.random.default_rng(2)
    n = 3000
    p = rng.uniform(-2, 2, (n, 3, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3, 3)).astype(np.float32)
    hits = 0
    with np.errstate(all="ignore"):
        for k in range(n):
            p0, p1, p2 = V(p[k, 0]), V(p[k, 1]), V(p[k, 2])
            target = (p[k, 0] * 0.3 + p[k, 1] * 0.3 + p[k, 2] * 0.4).astype(np.float32)
            if k % 5 == 0:
                target = p[k, k % 3]                         # through a vertex
            if k % 7 == 0:
                target = ((p[k, 0] + p[k, 1]) * np.float32(0.5)).astype(np.float32)   # through an edge
            x0 = (target + rng.normal(size=3).astype(np.float32) * 3).astype(np.float32)
            u = V(norm(V(target - x0))) if k % 3 else V(target - x0)      # also unnormalised directions
            flat = np.cross(p[k, 1] - p[k, 0], p[k, 2] - p[k, 0]).astype(np.float32)
            ns = [flat, flat, flat] if k % 2 else [nrm[k, 0], nrm[k, 1], nrm[k, 2]]
            if np.dot(np.asarray(u, np.float32), flat) > 0 and k % 4:
                ns = [-a for a in ns]                        # face the ray (the reference culls by shading normal)
            tmax = F(1e6 if k % 6 else 2.0)
            mine = collide_triangle(p0, p1, p2, V(ns[0]), V(ns[1]), V(ns[2]), V(x0), u, tmax)
            out = np.zeros(13, np.float32)
            got = L.yo_collide_triangle(ptr(arr(p0)), ptr(arr(p1)), ptr(arr(p2)), ptr(arr(ns[0])), ptr(arr(ns[1])),
                                        ptr(arr(ns[2])), ptr(arr(x0)), ptr(arr(u)), C.c_float(float(tmax)), ptr(out))
            assert bool(got) == (mine is not None), k
            if mine is not None:
                hits += 1
                t, x, n_, dpdu, dpdv = mine
                assert same_bits(out[0], t), k
                assert same_bits(out[1:4], x) and same_bits(out[4:7], n_) and same_bits(out[7:10], dpdu), k
                assert same_bits(out[10:13], dpdv), k
    assert hits > 300


def test_sphere_test_agrees_bit_for_bit(L):
    rng = np.random.default_rng(3)
    n = 3000
    hits = 0
    with np.errstate(all="ignore"):
        for k in range(n):
            s = rng.uniform(-3, 3, 3).astype(np.float32)
            r = F(rng.uniform(0.1, 2.0))
            x0 = rng.uniform(-6, 6, 3).astype(np.float32)
            if k % 4 == 0:
                x0 = (s + rng.normal(size=3).astype(np.float32) * np.float32(0.3)).astype(np.float32)   # inside
            aim = (s + rng.normal(size=3).astype(np.float32) * r * np.float32(0.9)).astype(np.float32)
            u = V(norm(V(aim - x0))) if k % 3 else V(aim - x0)
            if k % 10 == 0:                                   # straight at the pole: dpdu = n x (0,0,1) = 0
                x0 = (s + np.array([0, 0, 5], np.float32)).astype(np.float32)
                u = V((0, 0, -1))
            tmax = F(1e6 if k % 5 else 3.0)
            mine = collide_sphere(r, V(s), V(x0), u, tmax)
            out = np.zeros(13, np.float32)
            got = L.yo_collide_sphere(C.c_float(float(r)), ptr(arr(s)), ptr(arr(x0)), ptr(arr(u)), C.c_float(float(tmax)),
                                      ptr(out))
            assert bool(got) == (mine is not None), k
            if mine is not None:
                hits += 1
                t, x, n_, dpdu, dpdv = mine
                assert same_bits(out[0], t) and same_bits(out[1:4], x) and same_bits(out[4:7], n_), k
                assert same_bits(out[7:10], dpdu) and same_bits(out[10:13], dpdv), k
    assert hits > 1000


def test_bsdf_agrees(L):
    rng = np.random.default_rng(4)
    with np.errstate(all="ignore"):
        for k in range(2000):
            normal = rng.normal(size=3).astype(np.float32)           # NOT normalised (triangle shading normals)
            dpdu = np.cross(normal, rng.normal(size=3)).astype(np.float32)
            i_w = np.asarray(norm(V(rng.normal(size=3))), np.float32)
            o_w = np.asarray(norm(V(rng.normal(size=3))), np.float32)
            if k % 2:                                                # mostly above the horizon
                i_w = (i_w + normal / np.linalg.norm(normal)).astype(np.float32)
                o_w = (o_w + normal / np.linalg.norm(normal)).astype(np.float32)
            mat = np.concatenate([rng.uniform(0, 1, 6), [rng.choice([1.0, 10.0, 20.0, 200.0])]]).astype(np.float32)
            mine = bsdf_at(V(mat[0:3]), V(mat[3:6]), F(mat[6]), V(normal), V(dpdu), V(i_w), V(o_w))
            dg = arr(np.concatenate([np.zeros(3), normal, dpdu, np.zeros(3)]))
            out = np.zeros(3, np.float32)
            L.yo_bsdf_at(2, ptr(arr(mat)), ptr(dg), ptr(arr(i_w)), ptr(arr(o_w)), ptr(out))
            mine = np.asarray(mine, np.float32)
            assert np.array_equal(np.isnan(mine), np.isnan(out)), k
            ok = ~np.isnan(mine)
            assert np.array_equal(mine[ok] == 0, out[ok] == 0), k        # the horizon gate
            # identical arithmetic except powf (numpy's vs glibc's): a few ulp
            assert np.allclose(mine[ok], out[ok], rtol=4e-6, atol=0), (k, mine, out)


def test_camera_ray_agrees_bit_for_bit(L):
    rng = np.random.default_rng(5)
    for k in range(60):
        look = rng.normal(size=3).astype(np.float32)
        up = np.array([0, 1, 0], np.float32) if k % 2 else rng.normal(size=3).astype(np.float32)
        pos = rng.uniform(-50, 50, 3).astype(np.float32)
        imw, imh, focal = F(rng.choice([200, 512, 1920, 3840])), F(rng.choice([200, 389, 1080, 2160])), F(rng.choice([1, 1.5, 2]))
        cam = ob.make_camera(dict(imW=float(imw), imH=float(imh), focalLength=float(focal), lookDir=look.tolist(),
                                  upDir=up.tolist(), position=pos.tolist()))
        for (u, v) in [(0, 0), (int(imw) // 2, int(imh) // 2), (int(imw) - 1, int(imh) - 1), (17, 5)]:
            origin, d = camera_ray(imw, imh, focal, V(look), V(up), V(pos), u, v)
            out = np.zeros(7, np.float32)
            L.yo_camera_ray(C.byref(cam), C.c_float(u), C.c_float(v), ptr(out))
            assert same_bits(out[0:3], origin), (k, u, v)
            assert same_bits(out[3:6], d), (k, u, v)
            assert out[6] == np.float32(1e6)


# ---- Rays.hs:36-54, Lights.hs:15-24, Integrators.hs:22-61 -------------------------------------------------------
def collide_all(prims, x0, u, tmax):
    """collideAll over the primitive list in order: every Just replaces the hit and cuts the ray (Rays.hs:39-46).
    (The BVH returns the same hit as the plain list whenever no two hits tie exactly: Spec.hs:221-262.)"""
    best = None
    for k, pr in enumerate(prims):
        if pr[0] == "sphere":
            h = collide_sphere(pr[1], pr[2], x0, u, tmax)
        else:
            h = collide_triangle(*pr[1:7], x0, u, tmax)
        if h is not None:
            best = (k, h)
            tmax = h[0]
    return best


def reflection_dir(u, n):             # u - 2 * (u .* n) @* n                                   Integrators.hs:46-47
    return vsub(u, scale(F(2) * dot(u, n), n))


def radiance(prims, mats, lights, depth, x0, u):
    if depth == 0:
        return V((0, 0, 0))
    hit = collide_all(prims, x0, u, F(1e6))
    if hit is None:
        return V((0, 0, 0))
    k, (t, x, n, dpdu, _) = hit
    diffuse, specular, exponent = mats[prims[k][-1]]
    wo = (-u[0], -u[1], -u[2])
    r = reflection_dir(u, n)
    rs = radiance(prims, mats, lights, depth - 1, vadd(x, scale(F(0.001), r)), r)
    refl = vmul(scale(dot(n, r), bsdf_at(diffuse, specular, exponent, n, dpdu, r, wo)), rs)
    total = V((0, 0, 0))
    for (lpos, spectrum) in lights:
        ptl = vsub(lpos, x)
        ldir = norm(ptl)
        intensity = scale(F(1) / dot(ptl, ptl), spectrum)
        kk = bsdf_at(diffuse, specular, exponent, n, dpdu, ldir, wo)
        contrib = V((0, 0, 0))
        if dot(kk, kk) > 0:
            p0 = vadd(x, scale(F(0.001), ldir))
            d = vsub(lpos, p0)
            if collide_all(prims, p0, norm(d), np.sqrt(dot(d, d))) is None:          # reachable
                contrib = vmul(scale(abs(dot(ldir, n)), kk), intensity)
        total = vadd(total, contrib)
    return vadd(refl, total)


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_integrator_agrees(depth):
    """radiance / vhit / directIllumination / illuminationAtPoint / reachable, restated a second time, against the
    oracle on a small random scene (spheres + triangles, two lights); differences only through powf."""
    rng = np.random.default_rng(6)
    ns, nt = 6, 8
    sph_c = rng.uniform(-3, 3, (ns, 3)).astype(np.float32)
    sph_r = rng.uniform(0.4, 1.0, ns).astype(np.float32)
    tp = (rng.uniform(-4, 4, (nt, 1, 3)) + rng.uniform(-2.5, 2.5, (nt, 3, 3))).astype(np.float32)
    tn = np.cross(tp[:, 1] - tp[:, 0], tp[:, 2] - tp[:, 0]).astype(np.float32)
    toward_origin = (np.einsum("ij,ij->i", tn, -tp[:, 0]) < 0)
    tn[toward_origin] *= -1                                  # mostly face the camera region
    mats = np.array([[0.8, 0.6, 0.4, 0.3, 0.3, 0.3, 10], [0.2, 0.7, 0.9, 0.5, 0.5, 0.5, 40]], np.float32)
    lights = np.array([[5, 8, -6, 80, 80, 80], [-6, 4, 5, 40, 30, 20]], np.float32)
    sc = dict(tri_p0=tp[:, 0], tri_p1=tp[:, 1], tri_p2=tp[:, 2], tri_n0=tn, tri_n1=tn, tri_n2=tn,
              tri_material=(np.arange(nt) % 2).astype(np.uint32), sph_center=sph_c, sph_radius=sph_r,
              sph_material=(np.arange(ns) % 2).astype(np.uint32), prim_order=None, materials=mats, lights=lights,
              bvh_max_depth=16, split_mode=0)
    prims = [("sphere", F(sph_r[k]), V(sph_c[k]), k % 2) for k in range(ns)]
    prims += [("tri", V(tp[k, 0]), V(tp[k, 1]), V(tp[k, 2]), V(tn[k]), V(tn[k]), V(tn[k]), k % 2) for k in range(nt)]
    mm = [(V(m[0:3]), V(m[3:6]), F(m[6])) for m in mats]
    ll = [(V(li[0:3]), V(li[3:6])) for li in lights]
    o = ob.OracleScene(sc)
    lit = 0
    with np.errstate(all="ignore"):
        for k in range(150):
            x0 = rng.uniform(-9, 9, 3).astype(np.float32)
            aim = rng.uniform(-2.5, 2.5, 3).astype(np.float32)
            u = np.asarray(norm(V(aim - x0)), np.float32)
            mine = np.asarray(radiance(prims, mm, ll, depth, V(x0), V(u)), np.float32)
            got = o.radiance(x0, u, depth=depth)
            assert np.array_equal(np.isnan(mine), np.isnan(got)), k
            ok = ~np.isnan(mine)
            assert np.array_equal(mine[ok] == 0, np.asarray(got)[ok] == 0), (k, mine, got)
            assert np.allclose(mine[ok], np.asarray(got)[ok], rtol=2e-5, atol=1e-7), (k, mine, got)
            lit += int((mine[ok] > 0).any())
    o.close()
    assert lit > 40
