"""Deterministic synthetic scenes for the BASELINE.json configs.

Every generator returns `(scene, camera)`:
  scene  : dict of numpy arrays in the layout of `yahr_scene_desc` (include/yahr_b200.h) --
           primitives already expanded, in `objects >>= expand` order (main.hs:44), which IS the
           primitive ID.
  camera : dict with the fields of the reference's `Camera` record (Cameras.hs:54-56).

All geometry is float32.  Flat normals are computed exactly like `Scene.expand` does
(Scene.hs:78: n = norm ((p2 - p0) `cross` (p1 - p0))) with float32 arithmetic in the same order.
"""
import numpy as np

F = np.float32


# ------------------------------------------------------------------------------------------
# float32 helpers that follow Vectors.hs expression order
# ------------------------------------------------------------------------------------------
def _dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _norm(v):
    inv = F(1.0) / np.sqrt(_dot(v, v))
    return inv[..., None] * v


def flat_normals(p0, p1, p2):
    """Scene.hs:78 -- n = norm $ (p2 - p0) `cross` (p1 - p0), float32."""
    p0, p1, p2 = (np.asarray(p, F) for p in (p0, p1, p2))
    with np.errstate(all="ignore"):
        return _norm(_cross(p2 - p0, p1 - p0)).astype(F)


def _empty_scene():
    z3 = np.zeros((0, 3), F)
    return dict(tri_p0=z3, tri_p1=z3, tri_p2=z3, tri_n0=z3, tri_n1=z3, tri_n2=z3,
                tri_material=np.zeros(0, np.uint32), sph_center=z3, sph_radius=np.zeros(0, F),
                sph_material=np.zeros(0, np.uint32), prim_order=None,
                materials=np.zeros((0, 7), F), lights=np.zeros((0, 6), F), bvh_max_depth=16, split_mode=0,
                area_lights=np.zeros((0, 13), F))


def add_area_light(scene, corner, edge1, edge2, radiance, samples):
    """EXTENSION (no reference counterpart): appends a one-sided parallelogram light emitting towards
    edge1 x edge2 (row = corner, edge1, edge2, radiance, samples; include/yahr_b200.h)."""
    row = np.array([list(corner) + list(edge1) + list(edge2) + list(radiance) + [float(samples)]], F)
    prev = scene.get("area_lights")
    scene["area_lights"] = row if prev is None or len(prev) == 0 else np.concatenate([np.asarray(prev, F), row]).astype(F)
    return scene


def _camera(w, h, focal, look, up, pos):
    return dict(imW=float(w), imH=float(h), focalLength=float(focal), lookDir=[float(x) for x in look],
                upDir=[float(x) for x in up], position=[float(x) for x in pos])


def num_primitives(scene):
    return int(len(scene["tri_p0"]) + len(scene["sph_center"]))


def _floor(y, half=1000.0):
    """The 2-triangle floor of scene.yahrr:56-65 at height y (points and (0,1,2),(0,2,3))."""
    pts = np.array([[-half, y, -half], [half, y, -half], [half, y, half], [-half, y, half]], F)
    tris = np.array([[0, 1, 2], [0, 2, 3]])
    p0, p1, p2 = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    n = flat_normals(p0, p1, p2)
    return p0, p1, p2, n


# ------------------------------------------------------------------------------------------
# C1: the repo's own scene.yahrr, corrected to the current schema
# ------------------------------------------------------------------------------------------
def c1_scene_yahrr(width=512, height=512, recursion_native=False):
    """7 spheres + 2-triangle floor, 4 materials (scene.yahrr:9-65); light spectrum (500,500,500)
    is OUR choice (the file's light has none).  `recursion_native`: the file's own 1024x768."""
    sc = _empty_scene()
    sc["sph_center"] = np.array([[-2.4, -1.2, 20], [0, -1.2, 20], [2.4, -1.2, 20], [-1.2, -1.2, 15],
                                 [2, -1.2, 15], [0, -1.2, 10], [3, -1.2, 10]], F)
    sc["sph_radius"] = np.ones(7, F)
    sc["sph_material"] = np.zeros(7, np.uint32)           # "sphere"
    p0, p1, p2, n = _floor(-2.2)
    sc.update(tri_p0=p0, tri_p1=p1, tri_p2=p2, tri_n0=n, tri_n1=n, tri_n2=n,
              tri_material=np.full(2, 3, np.uint32))      # "bottom"
    # {diffuse3, specular3, shininess}; ambient is ignored by blinnPhong (Shaders.hs:12-14)
    sc["materials"] = np.array([[0.2, 0.2, 0.2, 1, 1, 1, 10],      # sphere
                                [0, 1, 0, 0, 0.1, 0, 1],            # back
                                [1, 0, 0, 0.1, 0, 0, 1],            # side
                                [1, 1, 1, 0.1, 0.1, 0.1, 1]], F)    # bottom
    sc["lights"] = np.array([[4, 10, -10, 500, 500, 500]], F)
    sc["bvh_max_depth"] = 16
    sc["split_mode"] = 0
    if recursion_native:
        width, height = 1024, 768
    cam = _camera(width, height, 1.5, [0.4, -0.3, 1], [0, 1, 0], [-4, 3, 2])
    return sc, cam


# ------------------------------------------------------------------------------------------
# value noise (deterministic, integer-hash lattice) for the bunny proxy and the terrain
# ------------------------------------------------------------------------------------------
def _hash3(ix, iy, iz, seed):
    h = (ix.astype(np.uint64) * np.uint64(0x9E3779B1) ^ iy.astype(np.uint64) * np.uint64(0x85EBCA77) ^
         iz.astype(np.uint64) * np.uint64(0xC2B2AE3D) ^ np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    return h.astype(np.float64) / 4294967296.0


def _value_noise3(p, seed):
    pf = np.floor(p)
    f = p - pf
    i = pf.astype(np.int64) & 0xFFFFF
    w = f * f * (3.0 - 2.0 * f)
    out = 0.0
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                c = _hash3(i[..., 0] + dx, i[..., 1] + dy, i[..., 2] + dz, seed)
                wx = w[..., 0] if dx else 1.0 - w[..., 0]
                wy = w[..., 1] if dy else 1.0 - w[..., 1]
                wz = w[..., 2] if dz else 1.0 - w[..., 2]
                out = out + c * wx * wy * wz
    return out


def _fbm3(p, seed, octaves=4):
    amp, freq, total, norm = 1.0, 1.0, 0.0, 0.0
    for o in range(octaves):
        total = total + amp * (_value_noise3(p * freq + 17.0 * o, seed + o) * 2.0 - 1.0)
        norm += amp
        amp *= 0.5
        freq *= 2.0
    return total / norm


def _smooth_vertex_normals(pts, tris):
    """Area-weighted vertex normals (float64 accumulate, float32 result)."""
    p = pts.astype(np.float64)
    fn = np.cross(p[tris[:, 1]] - p[tris[:, 0]], p[tris[:, 2]] - p[tris[:, 0]])
    vn = np.zeros_like(p)
    for k in range(3):
        np.add.at(vn, tris[:, k], fn)
    ln = np.linalg.norm(vn, axis=1, keepdims=True)
    ln[ln == 0] = 1.0
    return (vn / ln).astype(F)


def bunny_proxy_mesh(nu=187, nv=186, seed=0xB0221E, center=(0.0, 1.05, 0.0)):
    """Displaced UV sphere: nu x nv quads -> 2*nu*nv triangles, radius 1 + 0.15 * fbm, smooth
    vertex normals pointing outward.  Stand-in for the Stanford bunny, which is not in the
    reference repo and not obtainable offline."""
    theta = np.linspace(0.02 * np.pi, 0.98 * np.pi, nv + 1)        # tiny polar caps left open
    phi = np.arange(nu) * (2.0 * np.pi / nu)
    T, P = np.meshgrid(theta, phi, indexing="ij")                  # (nv+1, nu)
    d = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], axis=-1)
    r = 1.0 + 0.15 * _fbm3(d * 2.5 + 5.0, seed)
    pts = (d * r[..., None] + np.asarray(center)).reshape(-1, 3).astype(F)
    idx = np.arange((nv + 1) * nu).reshape(nv + 1, nu)
    a = idx[:-1, :]
    b = idx[1:, :]
    a2 = np.roll(a, -1, axis=1)
    b2 = np.roll(b, -1, axis=1)
    # winding chosen so that (p1-p0) x (p2-p0) points outward
    t1 = np.stack([a, a2, b], axis=-1).reshape(-1, 3)
    t2 = np.stack([a2, b2, b], axis=-1).reshape(-1, 3)
    tris = np.empty((t1.shape[0] * 2, 3), np.int64)
    tris[0::2] = t1
    tris[1::2] = t2
    vn = _smooth_vertex_normals(pts, tris)
    # make sure normals point away from the centre
    outward = ((pts - np.asarray(center, F)) * vn).sum(1) < 0
    vn[outward] *= -1
    return pts, tris, vn


def c2_bunny_proxy(width=1920, height=1080, nu=187, nv=186, bvh_depth=24, area_samples=0):
    """C2: bunny proxy (69 564 triangles at the default size) on the 2-triangle floor, exporter
    default material (render_engine.py:15-20), focalLength 2 (render_engine.py:157), one point light;
    area_samples > 0 adds the quad area light of the BASELINE config (extension) with that many samples."""
    pts, tris, vn = bunny_proxy_mesh(nu, nv)
    sc = _empty_scene()
    fp0, fp1, fp2, fn = _floor(0.0)
    sc["tri_p0"] = np.concatenate([pts[tris[:, 0]], fp0]).astype(F)
    sc["tri_p1"] = np.concatenate([pts[tris[:, 1]], fp1]).astype(F)
    sc["tri_p2"] = np.concatenate([pts[tris[:, 2]], fp2]).astype(F)
    sc["tri_n0"] = np.concatenate([vn[tris[:, 0]], fn]).astype(F)
    sc["tri_n1"] = np.concatenate([vn[tris[:, 1]], fn]).astype(F)
    sc["tri_n2"] = np.concatenate([vn[tris[:, 2]], fn]).astype(F)
    n = len(sc["tri_p0"])
    mat = np.zeros(n, np.uint32)
    mat[-2:] = 1
    sc["tri_material"] = mat
    sc["materials"] = np.array([[1, 1, 1, 0.3, 0.3, 0.3, 1],           # __default__ (exporter)
                                [0.8, 0.8, 0.8, 0.1, 0.1, 0.1, 20]], F)  # floor
    sc["lights"] = np.array([[4, 8, -6, 140, 140, 140]], F)
    if area_samples:      # a 2 x 2 panel above and to the left of the model, facing down
        add_area_light(sc, [-4.0, 5.0, -2.0], [2, 0, 0], [0, 0, 2], [12, 12, 12], area_samples)
    sc["bvh_max_depth"] = bvh_depth
    cam = _camera(width, height, 2.0, [-0.14, -0.12, 1.0], [0, 1, 0], [0.8, 1.7, -5.5])
    return sc, cam


# ------------------------------------------------------------------------------------------
# C3: sphere grid generalising Spec.hs:223-229
# ------------------------------------------------------------------------------------------
def c3_sphere_grid(n=32, width=2048, height=2048, bvh_depth=16):
    """Spheres of radius 1 at (10i, 10j, 10k), i,j,k in 1..n, listed i-major like the Spec's list
    comprehension; camera outside the grid looking at its centre; one point light above."""
    ax = np.arange(1, n + 1, dtype=F) * F(10)
    I, J, K = np.meshgrid(ax, ax, ax, indexing="ij")
    centers = np.stack([I, J, K], axis=-1).reshape(-1, 3).astype(F)
    sc = _empty_scene()
    sc["sph_center"] = centers
    sc["sph_radius"] = np.ones(len(centers), F)
    sc["sph_material"] = (np.arange(len(centers)) % 3).astype(np.uint32)
    sc["materials"] = np.array([[0.8, 0.2, 0.2, 0.5, 0.5, 0.5, 20], [0.2, 0.8, 0.2, 0.5, 0.5, 0.5, 5],
                                [0.2, 0.2, 0.8, 1, 1, 1, 50]], F)
    c = 5.0 * (n + 1)
    ext = 10.0 * n
    sc["lights"] = np.array([[c, c + 1.5 * ext, c - 0.3 * ext, 2.0 * ext * ext, 2.0 * ext * ext, 2.0 * ext * ext]], F)
    sc["bvh_max_depth"] = bvh_depth
    pos = np.array([c - 0.8 * ext, c + 0.7 * ext, c - 0.85 * ext])
    look = np.array([c, c, c]) - pos
    cam = _camera(width, height, 1.0, look, [0, 1, 0], pos)
    return sc, cam


# ------------------------------------------------------------------------------------------
# C4: 1M-triangle terrain and random soup
# ------------------------------------------------------------------------------------------
def c4_terrain(nx=1001, nz=501, width=3840, height=2160, seed=0x5EED, bvh_depth=32):
    """Height field of nx x nz vertices -> 2 (nx-1)(nz-1) triangles (1 000 000 at the default),
    4-octave value noise, smooth normals (pointing up)."""
    xs = np.linspace(-100.0, 100.0, nx)
    zs = np.linspace(-50.0, 50.0, nz)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    p = np.stack([X * 0.04, np.zeros_like(X), Z * 0.04], axis=-1)
    Y = 9.0 * _fbm3(p + 3.0, seed)
    pts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(F)
    idx = np.arange(nx * nz).reshape(nx, nz)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    t1 = np.stack([a, d, b], axis=-1).reshape(-1, 3)
    t2 = np.stack([b, d, c], axis=-1).reshape(-1, 3)
    tris = np.empty((t1.shape[0] * 2, 3), np.int64)
    tris[0::2] = t1
    tris[1::2] = t2
    vn = _smooth_vertex_normals(pts, tris)
    vn[vn[:, 1] < 0] *= -1
    sc = _empty_scene()
    for k, nm in enumerate(("tri_p0", "tri_p1", "tri_p2")):
        sc[nm] = pts[tris[:, k]]
    for k, nm in enumerate(("tri_n0", "tri_n1", "tri_n2")):
        sc[nm] = vn[tris[:, k]]
    sc["tri_material"] = np.zeros(len(tris), np.uint32)
    sc["materials"] = np.array([[0.6, 0.7, 0.5, 0.2, 0.2, 0.2, 8]], F)
    sc["lights"] = np.array([[60, 120, -30, 30000, 30000, 30000]], F)
    sc["bvh_max_depth"] = bvh_depth
    cam = _camera(width, height, 1.2, [0, -0.42, 1], [0, 1, 0], [0, 38, -95])
    return sc, cam


def c4_soup(n=1_000_000, width=3840, height=2160, seed=42, bvh_depth=32, half=50.0, edge=0.3):
    """Random triangle soup: centroids uniform in [-half, half]^3, edge vectors uniform in
    [-edge, edge]^3, flat normals (random facing; back-facing triangles never hit, Shapes.hs:57)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-half, half, (n, 3))
    e1 = rng.uniform(-edge, edge, (n, 3))
    e2 = rng.uniform(-edge, edge, (n, 3))
    p0 = (c - (e1 + e2) / 3.0).astype(F)
    p1 = (p0 + e1.astype(F)).astype(F)
    p2 = (p0 + e2.astype(F)).astype(F)
    nrm = flat_normals(p0, p1, p2)
    sc = _empty_scene()
    sc.update(tri_p0=p0, tri_p1=p1, tri_p2=p2, tri_n0=nrm, tri_n1=nrm, tri_n2=nrm,
              tri_material=(np.arange(n) % 2).astype(np.uint32))
    sc["materials"] = np.array([[0.9, 0.6, 0.3, 0.3, 0.3, 0.3, 10], [0.3, 0.6, 0.9, 0.3, 0.3, 0.3, 30]], F)
    sc["lights"] = np.array([[0, 3 * half, -3 * half, 40 * half * half, 40 * half * half, 40 * half * half]], F)
    sc["bvh_max_depth"] = bvh_depth
    cam = _camera(width, height, 1.0, [0, 0, 1], [0, 1, 0], [0, 0, -2.8 * half])
    return sc, cam


# ------------------------------------------------------------------------------------------
# C5: replicated bunny proxy (no instancing exists in Scene.hs, so real triangles)
# ------------------------------------------------------------------------------------------
def c5_replicated_bunny(copies=12, width=3840, height=2160, nu=187, nv=186, bvh_depth=40, spacing=3.0, area_samples=0):
    pts, tris, vn = bunny_proxy_mesh(nu, nv, center=(0.0, 1.05, 0.0))
    p = [pts[tris[:, k]] for k in range(3)]
    nn = [vn[tris[:, k]] for k in range(3)]
    outp = [[], [], []]
    outn = [[], [], []]
    for i in range(copies):
        for j in range(copies):
            off = np.array([(i - (copies - 1) / 2) * spacing, 0, j * spacing], F)
            for k in range(3):
                outp[k].append((p[k] + off).astype(F))
                outn[k].append(nn[k])
    fp0, fp1, fp2, fn = _floor(0.0)
    sc = _empty_scene()
    for k, nm in enumerate(("tri_p0", "tri_p1", "tri_p2")):
        sc[nm] = np.concatenate(outp[k] + [[fp0, fp1, fp2][k]]).astype(F)
    for k, nm in enumerate(("tri_n0", "tri_n1", "tri_n2")):
        sc[nm] = np.concatenate(outn[k] + [fn]).astype(F)
    n = len(sc["tri_p0"])
    mat = np.zeros(n, np.uint32)
    mat[-2:] = 1
    sc["tri_material"] = mat
    sc["materials"] = np.array([[1, 1, 1, 0.3, 0.3, 0.3, 1], [0.8, 0.8, 0.8, 0.1, 0.1, 0.1, 20]], F)
    span = copies * spacing
    sc["lights"] = np.array([[0.3 * span, 1.2 * span, -0.2 * span, 3 * span * span, 3 * span * span, 3 * span * span]], F)
    if area_samples:      # one large panel over the field, facing down (extension)
        add_area_light(sc, [-0.25 * span, 0.9 * span, 0.25 * span], [0.5 * span, 0, 0], [0, 0, 0.5 * span],
                       [6, 6, 6], area_samples)
    sc["bvh_max_depth"] = bvh_depth
    cam = _camera(width, height, 1.6, [0, -0.45, 1], [0, 1, 0], [0, 0.55 * span, -0.75 * span])
    return sc, cam


# ------------------------------------------------------------------------------------------
# adversarial scenes for the bit-exactness gate (SURVEY.md section 7 "hard parts")
# ------------------------------------------------------------------------------------------
def adversarial_shared_edges(width=256, height=256, n=24, bvh_depth=16, split_mode=0):
    """An axis-aligned camera staring at a regular triangulated wall at z = 10 whose vertices sit
    on exactly representable coordinates: many primary rays pass exactly through shared edges and
    vertices (equal-t ties between neighbours), have zero direction components, and graze flat
    (zero-thickness) boxes.  A second, coplanar copy of the wall makes every hit an exact tie."""
    xs = (np.arange(n + 1, dtype=F) - F(n / 2)) * F(0.5)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    pts = np.stack([X, Y, np.full_like(X, 10)], axis=-1).reshape(-1, 3).astype(F)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    t1 = np.stack([a, b, c], axis=-1).reshape(-1, 3)
    t2 = np.stack([a, c, d], axis=-1).reshape(-1, 3)
    tris = np.concatenate([t1, t2])
    p0, p1, p2 = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    nrm = np.tile(np.array([[0, 0, -1]], F), (len(tris), 1))
    sc = _empty_scene()
    sc.update(tri_p0=np.concatenate([p0, p0]), tri_p1=np.concatenate([p1, p1]), tri_p2=np.concatenate([p2, p2]),
              tri_n0=np.concatenate([nrm, nrm]), tri_n1=np.concatenate([nrm, nrm]), tri_n2=np.concatenate([nrm, nrm]),
              tri_material=np.concatenate([np.zeros(len(tris), np.uint32), np.ones(len(tris), np.uint32)]))
    # plus a few spheres tangent to the wall and to each other
    sc["sph_center"] = np.array([[0, 0, 9], [2, 0, 9], [0, 2, 9], [-2, -2, 8]], F)
    sc["sph_radius"] = np.array([1, 1, 1, 2], F)
    sc["sph_material"] = np.full(4, 2, np.uint32)
    # interleave spheres into the primitive order
    nt = 2 * len(tris)
    order = np.concatenate([(np.arange(nt, dtype=np.uint32) | np.uint32(1 << 31))[: nt // 2],
                            np.arange(4, dtype=np.uint32),
                            (np.arange(nt, dtype=np.uint32) | np.uint32(1 << 31))[nt // 2:]])
    sc["prim_order"] = order.astype(np.uint32)
    sc["materials"] = np.array([[1, 0.2, 0.2, 0.3, 0.3, 0.3, 4], [0.2, 1, 0.2, 0.3, 0.3, 0.3, 4],
                                [0.3, 0.3, 1, 1, 1, 1, 40]], F)
    sc["lights"] = np.array([[0, 0, 0, 60, 60, 60], [5, 5, 2, 30, 20, 10]], F)
    sc["bvh_max_depth"] = bvh_depth
    sc["split_mode"] = split_mode
    cam = _camera(width, height, 1.0, [0, 0, 1], [0, 1, 0], [0, 0, 0])
    return sc, cam


CONFIGS = {
    "c1": c1_scene_yahrr,
    "c2": c2_bunny_proxy,
    "c3": c3_sphere_grid,
    "c4-terrain": c4_terrain,
    "c4-soup": c4_soup,
    "c5": c5_replicated_bunny,
    "adversarial": adversarial_shared_edges,
}


# ------------------------------------------------------------------------------------------
# Degenerate inputs: what the reference does with them is defined by IEEE arithmetic alone
# ------------------------------------------------------------------------------------------
def degenerate_mix(width=160, height=120):
    """Zero-area and sliver triangles (1 / 0 and 0 * inf in the Moller-Trumbore quotient), a camera INSIDE a large sphere
    (second quadratic root), a small sphere whose +-z pole faces the camera (dpdu = n x (0,0,1) = 0: NaN shading frame,
    Shapes.hs:23-26), duplicated coplanar triangles (exact t ties), zero shading normals, a light sitting exactly on a
    surface, and a flat axis-aligned box in the BVH.  Geometry stays finite; the NaNs are produced by the arithmetic."""
    sc = _empty_scene()
    P = lambda *v: np.array(v, F)
    tris = [
        (P(-4, -1, 6), P(-2, -1, 6), P(-3, 1, 6)),          # regular, facing -z
        (P(-4, -1, 6), P(-2, -1, 6), P(-3, 1, 6)),          # exact duplicate: the later one wins the tie
        (P(2, 0, 5), P(2, 0, 5), P(2, 0, 5)),               # a point
        (P(-3, 0, 5), P(-2, 0, 5), P(-1, 0, 5)),            # collinear: zero area
        (P(-4, 2, 7), P(4, 2.0000002, 7), P(0, 2.0000001, 7)),   # sliver
        (P(-6, -2, 0), P(6, -2, 0), P(6, -2, 14)),          # floor (flat box in y)
        (P(-6, -2, 0), P(6, -2, 14), P(-6, -2, 14)),
        (P(1.5, -1, 4), P(2.5, -1, 4), P(2, 0, 4)),         # zero shading normals: u . ns = 0 is never < 0
    ]
    n = len(tris)
    p0 = np.stack([t[0] for t in tris]); p1 = np.stack([t[1] for t in tris]); p2 = np.stack([t[2] for t in tris])
    nn = flat_normals(p0, p1, p2)
    nn[np.isnan(nn)] = 0
    nn[0] = nn[1] = P(0, 0, -1)
    nn[5] = nn[6] = P(0, 1, 0)
    nn[7] = 0
    sc["tri_p0"], sc["tri_p1"], sc["tri_p2"] = p0, p1, p2
    sc["tri_n0"] = sc["tri_n1"] = sc["tri_n2"] = nn.astype(F)
    sc["tri_material"] = np.array([0, 1, 0, 0, 1, 2, 2, 0], np.uint32)
    sc["sph_center"] = np.array([[0, 0, 0], [0, 0, 9], [3, 1, 8]], F)    # the camera sits inside the first one
    sc["sph_radius"] = np.array([30, 0.5, 1.0], F)
    sc["sph_material"] = np.array([2, 1, 0], np.uint32)
    sc["materials"] = np.array([[0.9, 0.2, 0.2, 0.4, 0.4, 0.4, 8], [0.2, 0.9, 0.2, 0.2, 0.2, 0.2, 50],
                                [0.7, 0.7, 0.7, 0.0, 0.0, 0.0, 1]], F)
    sc["lights"] = np.array([[0, 4, 3, 60, 60, 60], [0, -2, 7, 20, 20, 20]], F)     # the second one lies ON the floor
    sc["bvh_max_depth"] = 16
    cam = _camera(width, height, 1.0, [0, 0, 1], [0, 1, 0], [0, 0, 0])
    return sc, cam
