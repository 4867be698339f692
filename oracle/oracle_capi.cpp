// oracle_capi.cpp -- C entry points of the CPU ORACLE (test infrastructure, see yahr_oracle.hpp).
//
// Used through ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  Never linked into libyahr_b200.so.
//
// The render loop follows main.hs:68-107: squareBatches tiles (Sampling.hs:5-21), each tile's
// pixels in u-major order, one `li (cast u v)` per pixel, scattered to index w*v+u.  Tiles are
// run as OpenMP dynamic tasks -- the analogue of renderPar (main.hs:86-96).
#include "yahr_oracle.hpp"

#include <chrono>
#include <cstdio>
#include <atomic>
#include <mutex>
#include <thread>
#include <sched.h>

using namespace yo;

extern "C" {

// Same field layout as include/yahr_b200.h's yahr_scene_desc / yahr_camera (declared
// independently here: the oracle shares no source with the product).
typedef struct {
  uint32_t n_triangles;
  const float *tri_p0, *tri_p1, *tri_p2, *tri_n0, *tri_n1, *tri_n2;
  const uint32_t* tri_material;
  uint32_t n_spheres;
  const float* sph_center;
  const float* sph_radius;
  const uint32_t* sph_material;
  const uint32_t* prim_order;  // NULL: spheres then triangles; else (kind<<31 | index) per primitive
  uint32_t n_materials;
  const float* materials;      // K x {diffuse3, specular3, shininess}
  uint32_t n_lights;
  const float* lights;         // L x {position3, spectrum3}
  int32_t bvh_max_depth;
  int32_t split_mode;          // 0 Midpoint, 1 SurfaceAreaHeuristic
  uint32_t n_area_lights;      // extension (no reference counterpart)
  const float* area_lights;    // A x {corner3, edge1 3, edge2 3, radiance3, samples}
} yo_scene_desc;

typedef struct {
  float imW, imH, focalLength;
  float lookDir[3], upDir[3], position[3];
} yo_camera;

typedef struct {
  uint64_t n_node, n_tri, n_tricand, n_sph, n_shade, n_light, n_primary, n_shadow, n_secondary;
  double seconds;       // wall time of the pixel loop
  int32_t threads;      // host threads used
  int32_t tiles;        // tiles rendered by this call
} yo_stats;

static Vec3 v3(const float* p) { return Vec3{p[0], p[1], p[2]}; }

void* yo_scene_create(const yo_scene_desc* d) {
  Scene* sc = new Scene();
  uint32_t n = d->n_spheres + d->n_triangles;
  sc->prims.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t ref = d->prim_order ? d->prim_order[i]
                                 : (i < d->n_spheres ? i : ((1u << 31) | (i - d->n_spheres)));
    uint32_t k = ref & 0x7fffffffu;
    Primitive& p = sc->prims[i];
    std::memset(&p, 0, sizeof(p));
    if (ref >> 31) {
      p.kind = 1;
      p.a = v3(d->tri_p0 + 3 * k); p.b = v3(d->tri_p1 + 3 * k); p.c = v3(d->tri_p2 + 3 * k);
      p.n0 = v3(d->tri_n0 + 3 * k); p.n1 = v3(d->tri_n1 + 3 * k); p.n2 = v3(d->tri_n2 + 3 * k);
      p.material = d->tri_material ? d->tri_material[k] : 0;
    } else {
      p.kind = 0;
      p.a = v3(d->sph_center + 3 * k);
      p.radius = d->sph_radius[k];
      p.material = d->sph_material ? d->sph_material[k] : 0;
    }
  }
  for (uint32_t i = 0; i < d->n_materials; ++i) {
    const float* m = d->materials + 7 * i;
    sc->materials.push_back(Material{v3(m), v3(m + 3), m[6]});
  }
  for (uint32_t i = 0; i < d->n_lights; ++i)
    sc->lights.push_back(Light{v3(d->lights + 6 * i), v3(d->lights + 6 * i + 3)});
  if (d->n_area_lights && d->area_lights)
    sc->areaLightData.assign(d->area_lights, d->area_lights + 13 * (size_t)d->n_area_lights);
  // buildCollisionModel (main.hs:41-53) + cull (main.hs:118)
  sc->bounds.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    const Primitive& p = sc->prims[i];
    sc->bounds[i] = p.kind == 0 ? boundSphere(p.radius, p.a) : boundTriangle(p.a, p.b, p.c);
  }
  BvhBuilder(sc->bounds, (SplitMode)d->split_mode, sc->bvh).build(d->bvh_max_depth);
  return sc;
}

void yo_scene_destroy(void* h) { delete (Scene*)h; }

// BVH inspection for the product's host-builder parity tests.
uint32_t yo_bvh_num_nodes(void* h) { return (uint32_t)((Scene*)h)->bvh.nodes.size(); }
uint32_t yo_bvh_num_prims(void* h) { return (uint32_t)((Scene*)h)->bvh.order.size(); }
int32_t yo_bvh_depth(void* h) { return ((Scene*)h)->bvh.maxDepthSeen; }
void yo_bvh_order(void* h, uint32_t* out) {
  Scene* sc = (Scene*)h;
  std::memcpy(out, sc->bvh.order.data(), sc->bvh.order.size() * sizeof(uint32_t));
}
// Pre-order (node, then left subtree, then right subtree) dump: per node
// kind(0 null,1 single,2 inner,3 multi), first, count, and the 6 box floats.
static void dumpNode(const Bvh& b, int32_t id, uint32_t* kinds, uint32_t* firsts, uint32_t* counts,
                     float* boxes, uint32_t& k) {
  const BvhNode& n = b.nodes[id];
  kinds[k] = n.kind; firsts[k] = n.first; counts[k] = n.count;
  float* bx = boxes + 6 * k;
  bx[0] = n.bb.bMin.x; bx[1] = n.bb.bMin.y; bx[2] = n.bb.bMin.z;
  bx[3] = n.bb.bMax.x; bx[4] = n.bb.bMax.y; bx[5] = n.bb.bMax.z;
  ++k;
  if (n.kind == BvhNode::Inner) {
    dumpNode(b, n.left, kinds, firsts, counts, boxes, k);
    dumpNode(b, n.right, kinds, firsts, counts, boxes, k);
  }
}
uint32_t yo_bvh_dump_preorder(void* h, uint32_t* kinds, uint32_t* firsts, uint32_t* counts, float* boxes) {
  Scene* sc = (Scene*)h;
  uint32_t k = 0;
  if (sc->bvh.root >= 0) dumpNode(sc->bvh, sc->bvh.root, kinds, firsts, counts, boxes, k);
  return k;
}

int64_t yo_num_batches(int64_t numThreads, int64_t w, int64_t h) { return numBatches(numThreads, w, h); }
void yo_batch_window(int64_t w, int64_t h, int64_t num, int64_t count, int32_t* out4) {
  Window win = batchWindow(w, h, num, count);
  out4[0] = win.x0; out4[1] = win.y0; out4[2] = win.x1; out4[3] = win.y1;
}

// Render tiles  tile_offset, tile_offset + tile_stride, ...  of the reference's own tiling.
// rgb: W*H*3 floats (row-major, RGB interleaved, row 0 = top; main.hs:98-107); primid: W*H
// uint32 (0xFFFFFFFF = miss) or NULL; tprim: W*H floats or NULL.  Pixels of tiles not
// rendered by this call are left untouched.  nthreads <= 0: all cores.
static int hostThreads() {
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) { const int n = CPU_COUNT(&set); if (n > 0) return n; }
  const unsigned hc = std::thread::hardware_concurrency();
  return hc ? (int)hc : 1;
}

int yo_render(void* h, const yo_camera* cam, int recursion_depth, int spp, uint64_t seed, float* rgb,
              uint32_t* primid, float* tprim, yo_stats* stats, int nthreads, int tile_stride,
              int tile_offset, int64_t n_batches_override) {
  Scene* sc = (Scene*)h;
  Camera c{cam->imW, cam->imH, cam->focalLength, v3(cam->lookDir), v3(cam->upDir), v3(cam->position)};
  Caster caster = makeCaster(c);
  const int64_t width = (int64_t)std::floor(c.imW), height = (int64_t)std::floor(c.imH);  // main.hs:122-123
  if (width <= 0 || height <= 0 || spp < 1 || tile_stride < 1) return 1;
  int threads = nthreads > 0 ? nthreads : hostThreads();
  const int64_t nBatches = n_batches_override > 0 ? n_batches_override : numBatches(1, width, height);
  Stats total;
  int tilesDone = 0;
  auto t0 = std::chrono::steady_clock::now();
  // One std::thread per host core; tiles are handed out through an atomic counter -- the analogue of renderPar's task
  // pool (main.hs:86-96).  Plain threads rather than OpenMP: with OMP_NUM_THREADS=1 exported (torchrun does that to
  // its workers) libgomp was seen to leave the whole loop to one thread whatever num_threads() said.
  std::atomic<int64_t> nextTile{0};
  std::mutex merge;
  auto worker = [&]() {
    Stats st;
    int myTiles = 0;
    for (;;) {
      const int64_t b = tile_offset + nextTile.fetch_add(1, std::memory_order_relaxed) * (int64_t)tile_stride;
      if (b >= nBatches) break;
      Window win = batchWindow(width, height, b, nBatches);
      ++myTiles;
      for (int u = win.x0; u < win.x1; ++u)                 // [(u, v) | u <- [x0..x1-1], v <- [y0..y1-1]]
        for (int v = win.y0; v < win.y1; ++v) {
          const uint32_t pixel = (uint32_t)(width * v + u); // uvToIndex (main.hs:100)
          Vec3 acc = vof(0.0f), L = vof(0.0f);
          int32_t prim0 = -1; float t0hit = 0.0f;
          for (int s = 0; s < spp; ++s) {
            float fu = (float)u + sampleOffset(seed, pixel, (uint32_t)s, 0);
            float fv = (float)v + sampleOffset(seed, pixel, (uint32_t)s, 1);
            Ray ray = computeInitialRay(caster, fu, fv);
            st.n_primary++;
            int32_t prim; float th;
            ShadeCtx ctx;
            ctx.seed = seed; ctx.pixel = pixel; ctx.sample = (uint32_t)s; ctx.level = 0;
            L = vcast(*sc, recursion_depth, ray, &st, &prim, &th, ctx);
            if (s == 0) { prim0 = prim; t0hit = th; }
            acc = vadd(acc, L);
          }
          Vec3 px = spp == 1 ? L : Vec3{acc.x / (float)spp, acc.y / (float)spp, acc.z / (float)spp};
          rgb[3 * (size_t)pixel + 0] = px.x; rgb[3 * (size_t)pixel + 1] = px.y; rgb[3 * (size_t)pixel + 2] = px.z;
          if (primid) primid[pixel] = (uint32_t)prim0;
          if (tprim) tprim[pixel] = t0hit;
        }
    }
    std::lock_guard<std::mutex> lock(merge);
    total.add(st);
    tilesDone += myTiles;
  };
  {
    std::vector<std::thread> pool;
    for (int k = 1; k < threads; ++k) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  if (stats) {
    stats->n_node = total.n_node; stats->n_tri = total.n_tri; stats->n_tricand = total.n_tricand;
    stats->n_sph = total.n_sph; stats->n_shade = total.n_shade; stats->n_light = total.n_light;
    stats->n_primary = total.n_primary; stats->n_shadow = total.n_shadow; stats->n_secondary = total.n_secondary;
    stats->seconds = std::chrono::duration<double>(t1 - t0).count();
    stats->threads = threads; stats->tiles = tilesDone;
  }
  return 0;
}

// ----------------------------------------------------------------------------------------
// Probes for the test/Spec.hs known-answer tests (tests/test_oracle_spec.py).
// ----------------------------------------------------------------------------------------
void yo_cross(const float* a, const float* b, float* out) {
  Vec3 r = cross(v3(a), v3(b)); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void yo_norm(const float* a, float* out) { Vec3 r = norm(v3(a)); out[0] = r.x; out[1] = r.y; out[2] = r.z; }

// collideAll over fake colliders `const (Just (Hit t_i undefined what_i))` / `const Nothing`
// (Spec.hs:50-60): present[i] != 0 means Just.  Returns `what` of the result or -1.
int yo_collide_all_const(int n, const int* present, const float* ts, const int* whats) {
  struct Fake { int present; float t; int what; };
  std::vector<Fake> fakes(n);
  for (int i = 0; i < n; ++i) fakes[i] = Fake{present[i], ts[i], whats[i]};
  Ray ray{vof(0.0f), vof(1.0f), 10.0f};
  auto r = collideAll(fakes.begin(), fakes.end(), ray, [](const Fake& f, const Ray&) -> std::optional<Hit> {
    if (!f.present) return std::nullopt;
    Hit h{}; h.t = f.t; h.what = f.what; return h;
  });
  return r ? r->what : -1;
}

// computeInitialRay: out = x0(3), u(3), tMax
void yo_camera_ray(const yo_camera* cam, float u, float v, float* out7) {
  Camera c{cam->imW, cam->imH, cam->focalLength, v3(cam->lookDir), v3(cam->upDir), v3(cam->position)};
  Ray r = computeInitialRay(makeCaster(c), u, v);
  out7[0] = r.x0.x; out7[1] = r.x0.y; out7[2] = r.x0.z; out7[3] = r.u.x; out7[4] = r.u.y; out7[5] = r.u.z;
  out7[6] = r.tMax;
}
void yo_camera_matrices(const yo_camera* cam, float* tf16, float* vtf16) {
  Camera c{cam->imW, cam->imH, cam->focalLength, v3(cam->lookDir), v3(cam->upDir), v3(cam->position)};
  Caster k = makeCaster(c);
  std::memcpy(tf16, k.tf.m, 64); std::memcpy(vtf16, k.vtf.m, 64);
}

static void putBox(const BoundingBox& b, float* out6) {
  out6[0] = b.bMin.x; out6[1] = b.bMin.y; out6[2] = b.bMin.z; out6[3] = b.bMax.x; out6[4] = b.bMax.y; out6[5] = b.bMax.z;
}
static BoundingBox getBox(const float* b) { return BoundingBox{v3(b), v3(b + 3)}; }
void yo_bound_sphere(float r, const float* s, float* out6) { putBox(boundSphere(r, v3(s)), out6); }
void yo_bound_triangle(const float* p0, const float* p1, const float* p2, float* out6) {
  putBox(boundTriangle(v3(p0), v3(p1), v3(p2)), out6);
}
void yo_bb_from_points(const float* a, const float* b, float* out6) { putBox(fromPoints(v3(a), v3(b)), out6); }
void yo_bb_include_point(const float* bb, const float* x, float* out6) { putBox(includePoint(getBox(bb), v3(x)), out6); }
void yo_bb_join(const float* a, const float* b, float* out6) { putBox(bbJoin(getBox(a), getBox(b)), out6); }
void yo_bb_centroid(const float* bb, float* out3) { Vec3 c = centroid(getBox(bb)); out3[0] = c.x; out3[1] = c.y; out3[2] = c.z; }
float yo_bb_surf(const float* bb) { return surf(getBox(bb)); }
// bbRayIntersection: returns 1 for Just, writes tNear.
int yo_bb_ray(const float* bb, const float* x0, const float* u, float tMax, float* tNear) {
  Ray r{v3(x0), v3(u), tMax};
  return bbRayIntersection(getBox(bb), r, tNear) ? 1 : 0;
}

// Shape probes: out = t, point(3), normal(3), dpdu(3), dpdv(3); returns 1 for Just.
static int putHit(const std::optional<Hit>& h, float* out13) {
  if (!h) return 0;
  out13[0] = h->t;
  const Vec3* v[4] = {&h->dg.dgPoint, &h->dg.dgNormal, &h->dg.dgDPDU, &h->dg.dgDPDV};
  for (int i = 0; i < 4; ++i) { out13[1 + 3 * i] = v[i]->x; out13[2 + 3 * i] = v[i]->y; out13[3 + 3 * i] = v[i]->z; }
  return 1;
}
int yo_collide_triangle(const float* p0, const float* p1, const float* p2, const float* n0, const float* n1,
                        const float* n2, const float* x0, const float* u, float tMax, float* out13) {
  Ray r{v3(x0), v3(u), tMax};
  return putHit(collideTriangle(0, v3(p0), v3(p1), v3(p2), v3(n0), v3(n1), v3(n2), r, nullptr), out13);
}
int yo_collide_sphere(float radius, const float* s, const float* x0, const float* u, float tMax, float* out13) {
  Ray r{v3(x0), v3(u), tMax};
  return putHit(collideSphere(0, radius, v3(s), r, nullptr), out13);
}

// Root collider probe on a built scene: returns primitive ID or -1; out13 as above.
int yo_scene_collide(void* h, const float* x0, const float* u, float tMax, float* out13) {
  Scene* sc = (Scene*)h;
  Ray r{v3(x0), v3(u), tMax};
  auto hit = rootCollider(*sc, r, nullptr);
  if (!hit) return -1;
  if (out13) putHit(hit, out13);
  return hit->what;
}
// radiance probe on a built scene for an arbitrary ray (Integrators.hs:22).
void yo_scene_radiance(void* h, int recursion_depth, const float* x0, const float* u, float tMax, float* out3) {
  Scene* sc = (Scene*)h;
  Ray r{v3(x0), v3(u), tMax};
  Vec3 L = vcast(*sc, recursion_depth, r, nullptr);
  out3[0] = L.x; out3[1] = L.y; out3[2] = L.z;
}
// BSDF.at probes.  kind 0: Lambertian, 1: Blinn exponent, 2: blinnPhong composite(material7).
void yo_bsdf_at(int kind, const float* params, const float* dg12, const float* i, const float* o, float* out3) {
  DifferentialGeometry dg{v3(dg12), v3(dg12 + 3), v3(dg12 + 6), v3(dg12 + 9)};
  Material m;
  if (kind == 0) m = Material{vof(1.0f), vof(0.0f), 1.0f};
  else if (kind == 1) m = Material{vof(0.0f), vof(1.0f), params[0]};
  else m = Material{v3(params), v3(params + 3), params[6]};
  Vec3 r = bsdfAt(m, dg, v3(i), v3(o));
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
float yo_sample_offset(uint64_t seed, uint32_t pixel, uint32_t s, uint32_t dim) { return sampleOffset(seed, pixel, s, dim); }
int yo_max_threads(void) { return hostThreads(); }

}  // extern "C"
