// yahr_oracle.hpp -- CPU ORACLE. TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A binary32, FMA-free restatement of the per-pixel render loop of pstiasny/yahr
// (a pure-Haskell Whitted ray tracer).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may link or call this.  The product
// (yahr_b200/csrc, libyahr_b200.so) never includes this file.
//
// PARITY PINNING: the reference cannot be compiled here (no GHC, no Hackage packages), so
// this restatement is pinned against every known-answer / property test in the reference's
// own test/Spec.hs (see tests/test_oracle_spec.py).  Those are eps=1e-4 or ordering
// assertions; no reference artefact pins an image or a radiance value, so image-level
// parity is "parity unpinned" beyond those KATs.
//
// Must be compiled with:  -O2 -ffp-contract=off   and WITHOUT -ffast-math.
// Every function cites the reference file:line it follows (paths relative to the
// reference repository root).  Expression ORDER is part of the contract: Haskell's
// infixl 6/7 parse is written out with explicit parentheses.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <optional>
#include <vector>

#if defined(__FAST_MATH__)
#error "the oracle must not be compiled with -ffast-math"
#endif

namespace yo {

// ---------------------------------------------------------------------------------------
// Vectors.hs
// ---------------------------------------------------------------------------------------
struct Vec3 { float x, y, z; };                        // Vectors.hs:5  (three strict Floats)

inline Vec3 vof(float l) { return {l, l, l}; }         // Vectors.hs:20-21
inline Vec3 vadd(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }   // :24
inline Vec3 vsub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }   // :25
inline Vec3 vmul(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }   // :26
inline Vec3 vneg(Vec3 a) { return {-a.x, -a.y, -a.z}; }                          // :27
inline Vec3 vabs(Vec3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; } // :28

// (.*), infixl 7:  x1*y1 + x2*y2 + x3*y3  ==  (x1*y1 + x2*y2) + x3*y3      Vectors.hs:32-34
inline float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float lensq(Vec3 v) { return dot(v, v); }                                 // :36-37
inline float len(Vec3 v) { return std::sqrt(lensq(v)); }                         // :39-40
inline Vec3 scale(float f, Vec3 v) { return vmul(vof(f), v); }                   // (@*) :42-44
inline Vec3 norm(Vec3 v) { return scale(1.0f / len(v), v); }                     // :46-47
inline Vec3 cross(Vec3 a, Vec3 b) {                                              // :49-53
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

enum Dimension { X = 0, Y = 1, Z = 2 };                                          // :55
inline float getDimension(Dimension d, Vec3 v) { return d == X ? v.x : (d == Y ? v.y : v.z); } // :57-60
inline Dimension maxDimension(Vec3 v) {                                          // :62-66
  if (v.x > v.y && v.x > v.z) return X;
  if (v.y > v.z) return Y;
  return Z;
}

// GHC's `Ord Float` instance defines only the comparisons; min/max are the class defaults
//   max x y = if x <= y then y else x ;  min x y = if x <= y then x else y
// which are NOT fminf/fmaxf when a NaN is involved (SURVEY.md note N).
inline float hs_min(float x, float y) { return x <= y ? x : y; }
inline float hs_max(float x, float y) { return x <= y ? y : x; }
inline Vec3 vzip_min(Vec3 a, Vec3 b) { return {hs_min(a.x, b.x), hs_min(a.y, b.y), hs_min(a.z, b.z)}; }
inline Vec3 vzip_max(Vec3 a, Vec3 b) { return {hs_max(a.x, b.x), hs_max(a.y, b.y), hs_max(a.z, b.z)}; }

// ---------------------------------------------------------------------------------------
// DifferentialGeometry.hs:5-10, Rays.hs
// ---------------------------------------------------------------------------------------
struct DifferentialGeometry { Vec3 dgPoint, dgNormal, dgDPDU, dgDPDV; };

struct Ray { Vec3 x0, u; float tMax; };                                          // Rays.hs:11

// Hit t dg a (Rays.hs:16).  `what` is our primitive ID: the index of the primitive in
// `objects >>= expand` order (main.hs:44); the material is looked up through it.
struct Hit { float t; DifferentialGeometry dg; int32_t what; };

inline Vec3 rayAt(const Ray& r, float t) { return vadd(r.x0, scale(t, r.u)); }   // Rays.hs:35

// Work counters.  These define the "algorithmic bytes per ray" of SURVEY.md 8(d).
struct Stats {
  uint64_t n_node = 0;      // boxes tested (bbRayIntersection calls)
  uint64_t n_tri = 0;       // collideTriangle calls
  uint64_t n_tricand = 0;   // triangle tests that reach the shading-normal test
  uint64_t n_sph = 0;       // collideSphere calls
  uint64_t n_shade = 0;     // material evaluations (vhit)
  uint64_t n_light = 0;     // light samples (illuminationAtPoint)
  uint64_t n_primary = 0;   // camera rays
  uint64_t n_shadow = 0;    // shadow probes actually traced (Integrators.hs:59 short circuit)
  uint64_t n_secondary = 0; // reflection rays (recursionDepth > 1)
  void add(const Stats& o) {
    n_node += o.n_node; n_tri += o.n_tri; n_tricand += o.n_tricand; n_sph += o.n_sph;
    n_shade += o.n_shade; n_light += o.n_light; n_primary += o.n_primary;
    n_shadow += o.n_shadow; n_secondary += o.n_secondary;
  }
};

// collideAll (Rays.hs:39-46): left fold; EVERY `Just h` replaces the current hit and cuts
// the ray to tHit h.  Correctness of "nearest" relies on each collider enforcing t <= tMax.
template <class It, class Call>
inline std::optional<Hit> collideAll(It first, It last, const Ray& ray0, Call&& call) {
  std::optional<Hit> mhit;                        // (Nothing, ray)
  Ray ray = ray0;
  for (It it = first; it != last; ++it) {
    std::optional<Hit> h = call(*it, ray);        // case collider ray of
    if (h) { mhit = h; ray.tMax = h->t; }         //   Just h -> (Just h, cutRay (tHit h) ray)
  }                                               //   Nothing -> (mhit, ray)
  return mhit;
}

// ---------------------------------------------------------------------------------------
// AABBs.hs
// ---------------------------------------------------------------------------------------
struct BoundingBox { Vec3 bMin, bMax; };                                         // AABBs.hs:8

inline BoundingBox bbEmpty() {                                                   // :10-11
  const float inf = std::numeric_limits<float>::infinity();
  return {vof(inf), vof(-inf)};
}
inline BoundingBox fromPoints(Vec3 x, Vec3 y) { return {vzip_min(x, y), vzip_max(x, y)}; }  // :13-14
inline BoundingBox includePoint(const BoundingBox& b, Vec3 x) {                  // :16-18
  return {vzip_min(b.bMin, x), vzip_max(b.bMax, x)};
}
inline BoundingBox bbJoin(const BoundingBox& a, const BoundingBox& b) {          // :25-27
  return {vzip_min(a.bMin, b.bMin), vzip_max(a.bMax, b.bMax)};
}

// bbRayIntersection (AABBs.hs:29-40).  Returns true for `Just tNear`.  `tNearOut` optional.
inline bool bbRayIntersection(const BoundingBox& bb, const Ray& r, float* tNearOut = nullptr) {
  float tNear = 0.0f, tFar = r.tMax;                       // foldl ... (0, tMax) [X, Y, Z]
  for (int d = 0; d < 3; ++d) {
    Dimension dim = (Dimension)d;
    float invRayDir = 1.0f / getDimension(dim, r.u);
    float t0 = (getDimension(dim, bb.bMin) - getDimension(dim, r.x0)) * invRayDir;
    float t1 = (getDimension(dim, bb.bMax) - getDimension(dim, r.x0)) * invRayDir;
    float tDimNear = hs_min(t0, t1);
    float tDimFar = hs_max(t0, t1);
    tNear = hs_max(tNear, tDimNear);
    tFar = hs_min(tFar, tDimFar);
  }
  if (tNearOut) *tNearOut = tNear;
  return !(tNear > tFar);                                  // if tNear > tFar then Nothing
}

inline Dimension maxExtent(const BoundingBox& b) { return maxDimension(vsub(b.bMax, b.bMin)); } // :45-46
inline Vec3 centroid(const BoundingBox& b) {                                     // :48-49
  return vadd(scale(0.5f, b.bMin), scale(0.5f, b.bMax));
}
inline float surf(const BoundingBox& b) {                                        // :51-53
  Vec3 d = vsub(b.bMax, b.bMin);
  return 2.0f * ((d.x * d.y + d.x * d.z) + d.y * d.z);
}

// boundAllPoints (AABBs.hs:20-23); caller guarantees n >= 1 (the reference `error`s on []).
template <class GetPoint>
inline BoundingBox boundAllPoints(size_t n, GetPoint&& pt) {
  if (n == 1) return fromPoints(pt(0), pt(0));
  BoundingBox b = fromPoints(pt(0), pt(1));
  for (size_t i = 2; i < n; ++i) b = includePoint(b, pt(i));
  return b;
}

// ---------------------------------------------------------------------------------------
// Shapes.hs
// ---------------------------------------------------------------------------------------
// collideSphere (Shapes.hs:13-29)
inline std::optional<Hit> collideSphere(int32_t what, float r, Vec3 s, const Ray& ray, Stats* st) {
  if (st) st->n_sph++;
  const Vec3 x0 = ray.x0, u = ray.u;
  Vec3 d = vsub(x0, s);
  float a = dot(u, u);
  float b = dot(vmul(vof(2.0f), d), u);        // 2 * d .* u  ==  (2 * d) .* u   (infixl 7)
  float c = dot(d, d) - r * r;
  float delta = b * b - (4.0f * a) * c;        // b * b - 4 * a * c
  if (delta < 0) return std::nullopt;
  float ts[2] = {(-b - std::sqrt(delta)) / (2.0f * a), (-b + std::sqrt(delta)) / (2.0f * a)};
  for (int i = 0; i < 2; ++i) {                // listToMaybe [ ... | t <- ts, t > 0, t <= tMax ]
    float t = ts[i];
    if (t > 0 && t <= ray.tMax) {
      Vec3 x = rayAt(ray, t);
      Vec3 n = norm(vsub(x, s));
      Hit h;
      h.t = t;
      h.dg.dgPoint = x;
      h.dg.dgNormal = n;
      h.dg.dgDPDU = cross(n, Vec3{0, 0, 1});
      h.dg.dgDPDV = cross(n, Vec3{0, 1, 0});
      h.what = what;
      return h;
    }
  }
  return std::nullopt;
}

inline BoundingBox boundSphere(float r, Vec3 s) {                                // Shapes.hs:32-33
  return fromPoints(vadd(s, vof(r)), vsub(s, vof(r)));
}

// collideTriangle (Shapes.hs:36-59)
inline std::optional<Hit> collideTriangle(int32_t what, Vec3 p0, Vec3 p1, Vec3 p2, Vec3 n0, Vec3 n1,
                                          Vec3 n2, const Ray& ray, Stats* st) {
  if (st) st->n_tri++;
  const Vec3 x0 = ray.x0, u = ray.u;
  Vec3 e1 = vsub(p1, p0);
  Vec3 e2 = vsub(p2, p0);
  Vec3 s = vsub(x0, p0);
  Vec3 s1 = cross(u, e2);
  Vec3 s2 = cross(s, e1);
  float invDiv = 1.0f / dot(s1, e1);
  float t = dot(s2, e2) * invDiv;
  float b1 = dot(s1, s) * invDiv;
  float b2 = dot(s2, u) * invDiv;
  float b0 = (1.0f - b1) - b2;
  if (!(b0 >= 0 && b0 <= 1 && b1 >= 0 && b1 <= 1 && b2 >= 0 && b2 <= 1 && t > 0 && t <= ray.tMax))
    return std::nullopt;
  if (st) st->n_tricand++;
  Vec3 ns = vadd(vadd(scale(b0, n0), scale(b1, n1)), scale(b2, n2));
  if (!(dot(u, ns) < 0)) return std::nullopt;
  Vec3 ss = norm(e2);
  Vec3 ts = norm(cross(ss, ns));
  Hit h;
  h.t = t;
  h.dg.dgPoint = vadd(x0, scale(t, u));
  h.dg.dgNormal = ns;
  h.dg.dgDPDU = cross(ts, ns);
  h.dg.dgDPDV = ts;
  h.what = what;
  return h;
}

inline BoundingBox boundTriangle(Vec3 p0, Vec3 p1, Vec3 p2) {                    // Shapes.hs:62-63
  return includePoint(fromPoints(p0, p1), p2);
}

// ---------------------------------------------------------------------------------------
// Scene-side data (main.hs:41-60): the expanded object list, materials, lights.
// ---------------------------------------------------------------------------------------
struct Primitive {
  uint8_t kind;           // 0 = Sphere, 1 = Triangle (Scene.hs:15-29 after expand)
  Vec3 a, b, c;           // sphere: a = position; triangle: p0 p1 p2
  Vec3 n0, n1, n2;        // triangle vertex normals
  float radius;           // sphere
  uint32_t material;      // index into materials (main.hs:55 `mats ! mId`)
};
struct Material { Vec3 diffuse, specular; float shininess; };   // Scene.hs:45-50; ambient is ignored (Shaders.hs:12-14)
struct Light { Vec3 position, spectrum; };                       // Lights.hs:7
struct AreaLight;                                                // extension, defined with the lights code below

// ---------------------------------------------------------------------------------------
// Culling.hs -- the closure tree reified as an explicit node tree.
// ---------------------------------------------------------------------------------------
enum SplitMode { Midpoint = 0, SurfaceAreaHeuristic = 1 };       // Culling.hs:18

struct BvhNode {
  enum Kind : uint8_t { Null, Single, Inner, Multi } kind = Null;
  BoundingBox bb;                 // the box handed to wrapCollider
  int32_t left = -1, right = -1;  // Inner
  uint32_t first = 0, count = 0;  // Single/Multi: range in Bvh::order
};

struct Bvh {
  std::vector<BvhNode> nodes;
  std::vector<uint32_t> order;    // primitive IDs in left-first DFS leaf order
  int32_t root = -1;              // -1: `bvh _ _ [] = const Nothing` (Culling.hs:28)
  int maxDepthSeen = 0;
};

struct Scene {
  std::vector<Primitive> prims;
  std::vector<Material> materials;
  std::vector<Light> lights;
  std::vector<float> areaLightData;  // extension: 13 floats per area light (corner, edge1, edge2, radiance, samples)
  std::vector<BoundingBox> bounds;   // boundSceneObject, main.hs:51-53
  Bvh bvh;
};

class BvhBuilder {
 public:
  BvhBuilder(const std::vector<BoundingBox>& bbs, SplitMode mode, Bvh& out) : bbs_(bbs), mode_(mode), out_(out) {}

  // bvh (Culling.hs:27-29)
  void build(int maxDepth) {
    out_.nodes.clear(); out_.order.clear(); out_.root = -1; out_.maxDepthSeen = 0;
    if (bbs_.empty()) return;
    std::vector<uint32_t> xs(bbs_.size());
    for (size_t i = 0; i < xs.size(); ++i) xs[i] = (uint32_t)i;
    out_.root = buildTree(maxDepth, xs, 0);
  }

 private:
  const std::vector<BoundingBox>& bbs_;
  SplitMode mode_;
  Bvh& out_;

  int32_t newNode() { out_.nodes.emplace_back(); return (int32_t)out_.nodes.size() - 1; }

  // multiLeaf (Culling.hs:52-54): jointBB = foldl join empty bbs
  int32_t multiLeaf(const std::vector<uint32_t>& xs) {
    int32_t id = newNode();
    BoundingBox jb = bbEmpty();
    for (uint32_t i : xs) jb = bbJoin(jb, bbs_[i]);
    BvhNode& n = out_.nodes[id];
    n.kind = BvhNode::Multi; n.bb = jb; n.first = (uint32_t)out_.order.size(); n.count = (uint32_t)xs.size();
    out_.order.insert(out_.order.end(), xs.begin(), xs.end());
    return id;
  }

  // buildTree (Culling.hs:31-50).  Children are built left first so that Bvh::order is the
  // left-first DFS leaf order (the order `collideAll` visits primitives in).
  int32_t buildTree(int maxDepth, const std::vector<uint32_t>& xs, int level) {
    if (level > out_.maxDepthSeen) out_.maxDepthSeen = level;
    if (xs.empty()) {                                        // buildTree _ [] = (empty, nullCollider)
      int32_t id = newNode();
      out_.nodes[id].kind = BvhNode::Null; out_.nodes[id].bb = bbEmpty();
      return id;
    }
    if (xs.size() == 1) {                                    // buildTree _ [(bb, col)] = (bb, wrapCollider col bb)
      int32_t id = newNode();
      BvhNode& n = out_.nodes[id];
      n.kind = BvhNode::Single; n.bb = bbs_[xs[0]]; n.first = (uint32_t)out_.order.size(); n.count = 1;
      out_.order.push_back(xs[0]);
      return id;
    }
    if (maxDepth == 0) return multiLeaf(xs);                 // buildTree 0 xs = multiLeaf xs

    // bbOfCentroids = boundAllPoints (map centroid bbs) ; dim = maxExtent bbOfCentroids
    BoundingBox bbOfCentroids = boundAllPoints(xs.size(), [&](size_t i) { return centroid(bbs_[xs[i]]); });
    Dimension dim = maxExtent(bbOfCentroids);
    std::vector<uint32_t> xs1, xs2;
    if (mode_ == SurfaceAreaHeuristic) sahSplit(xs, dim, bbOfCentroids, xs1, xs2);
    else midpointSplit(xs, dim, bbOfCentroids, xs1, xs2);

    if (xs2.empty()) return multiLeaf(xs1);                  // if null xs2 then multiLeaf xs1
    int32_t id = newNode();
    int32_t l = buildTree(maxDepth - 1, xs1, level + 1);
    int32_t r = buildTree(maxDepth - 1, xs2, level + 1);
    BvhNode& n = out_.nodes[id];
    n.kind = BvhNode::Inner; n.left = l; n.right = r;
    n.bb = bbJoin(out_.nodes[l].bb, out_.nodes[r].bb);       // jointBB = join bb1 bb2
    return id;
  }

  // midpointSplit (Culling.hs:56-60): stable `partition (<= midpoint)` on the centroid coordinate
  void midpointSplit(const std::vector<uint32_t>& xs, Dimension dim, const BoundingBox& bbOfCentroids,
                     std::vector<uint32_t>& a, std::vector<uint32_t>& b) {
    float midpoint = getDimension(dim, centroid(bbOfCentroids));
    for (uint32_t i : xs) {
      if (getDimension(dim, centroid(bbs_[i])) <= midpoint) a.push_back(i); else b.push_back(i);
    }
  }

  // Haskell `floor x :: Int` for a Float through float2Int#: on x86-64 this is cvttss2si of the
  // floored value, which yields INT64_MIN for NaN / out-of-range inputs.
  static int64_t hsFloorToInt(float x) {
    float f = std::floor(x);
    if (!(f >= -9.2233720368547758e18f && f < 9.2233720368547758e18f)) return INT64_MIN;
    return (int64_t)f;
  }

  // sahSplit (Culling.hs:62-112), quirks included: wrappedArea is the SUM of per-bucket surface
  // areas and surf empty = +inf, so any empty bucket collapses every cost to 0.125 or NaN.
  void sahSplit(const std::vector<uint32_t>& xs, Dimension dim, const BoundingBox& bbOfCentroids,
                std::vector<uint32_t>& a, std::vector<uint32_t>& b) {
    const int nBuckets = 16;
    float centMin = getDimension(dim, bbOfCentroids.bMin);
    float centMax = getDimension(dim, bbOfCentroids.bMax);
    auto bucketId = [&](const BoundingBox& bb) -> int {
      float lenv = centMax - centMin;
      float d = getDimension(dim, centroid(bb)) - centMin;
      float frac = d / lenv;
      int64_t fl = hsFloorToInt((float)nBuckets * frac);
      int64_t m = fl > 0 ? fl : 0;                           // max 0
      return (int)(m < nBuckets - 1 ? m : nBuckets - 1);     // min (nBuckets - 1)
    };
    if (!(xs.size() > 4)) { midpointSplit(xs, dim, bbOfCentroids, a, b); return; }  // :110-112

    int counts[16]; BoundingBox boxes[16];
    for (int i = 0; i < nBuckets; ++i) { counts[i] = 0; boxes[i] = bbEmpty(); }
    for (uint32_t i : xs) { int k = bucketId(bbs_[i]); counts[k]++; boxes[k] = bbJoin(boxes[k], bbs_[i]); }

    float wrappedArea = 0.0f;                                // foldl (\x (_, bb) -> x + surf bb) 0 buckets
    for (int i = 0; i < nBuckets; ++i) wrappedArea = wrappedArea + surf(boxes[i]);

    auto costOfSplitAt = [&](int i) -> float {
      BoundingBox b0 = bbEmpty(), b1 = bbEmpty();
      int c0 = 0, c1 = 0;
      for (int k = 0; k <= i; ++k) { b0 = bbJoin(b0, boxes[k]); c0 += counts[k]; }
      for (int k = i + 1; k < nBuckets; ++k) { b1 = bbJoin(b1, boxes[k]); c1 += counts[k]; }
      float count0 = (float)c0, count1 = (float)c1;
      return 0.125f + (count0 * surf(b0) + count1 * surf(b1)) / wrappedArea;
    };
    // minimumBy (compare `on` snd): acc is replaced iff `compare acc next == GT`, where
    // compare x y = if x < y then LT else if x == y then EQ else GT  (so any NaN => GT).
    int minSplit = 0; float minCost = costOfSplitAt(0);
    for (int i = 1; i <= nBuckets - 2; ++i) {
      float c = costOfSplitAt(i);
      bool keep = (minCost < c) || (minCost == c);
      if (!keep) { minSplit = i; minCost = c; }
    }
    for (uint32_t i : xs) { if (bucketId(bbs_[i]) <= minSplit) a.push_back(i); else b.push_back(i); }
  }
};

// The root collider (Culling.hs:24-25,33,38,52) evaluated on the reified tree.
inline std::optional<Hit> collidePrimitive(const Scene& sc, uint32_t id, const Ray& ray, Stats* st) {
  const Primitive& p = sc.prims[id];
  if (p.kind == 0) return collideSphere((int32_t)id, p.radius, p.a, ray, st);      // main.hs:47
  return collideTriangle((int32_t)id, p.a, p.b, p.c, p.n0, p.n1, p.n2, ray, st);   // main.hs:48-49
}

inline std::optional<Hit> collideNode(const Scene& sc, int32_t node, const Ray& ray, Stats* st) {
  const BvhNode& n = sc.bvh.nodes[node];
  if (n.kind == BvhNode::Null) return std::nullopt;          // nullCollider (not wrapped, Culling.hs:32)
  // wrapCollider cf bb r = bbRayIntersection bb r >>= \t -> cf r        (AABBs.hs:42-43)
  if (st) st->n_node++;
  if (!bbRayIntersection(n.bb, ray)) return std::nullopt;
  if (n.kind == BvhNode::Single) return collidePrimitive(sc, sc.bvh.order[n.first], ray, st);
  if (n.kind == BvhNode::Inner) {                            // collideBVHNodes: collideAll [col1, col2]
    const int32_t kids[2] = {n.left, n.right};
    return collideAll(kids, kids + 2, ray, [&](int32_t k, const Ray& r) { return collideNode(sc, k, r, st); });
  }
  const uint32_t* first = sc.bvh.order.data() + n.first;     // multiLeaf: collideAll colliders
  return collideAll(first, first + n.count, ray,
                    [&](uint32_t id, const Ray& r) { return collidePrimitive(sc, id, r, st); });
}

inline std::optional<Hit> rootCollider(const Scene& sc, const Ray& ray, Stats* st) {
  if (sc.bvh.root < 0) return std::nullopt;
  return collideNode(sc, sc.bvh.root, ray, st);
}

// reachable (Rays.hs:49-54): a FULL closest-hit probe of which only emptiness is used.
inline bool reachable(const Scene& sc, Vec3 p0, Vec3 p1, Stats* st) {
  Ray probe{p0, norm(vsub(p1, p0)), len(vsub(p1, p0))};
  return !rootCollider(sc, probe, st).has_value();
}

// ---------------------------------------------------------------------------------------
// BSDF.hs, Shaders.hs
// ---------------------------------------------------------------------------------------
static const float kPi = 3.14159274101257324f;              // pi :: Float = 0x40490FDB

// locAt Lambertian (BSDF.hs:15-17)
inline Vec3 locAtLambertian(Vec3 i, Vec3 o) {
  if (i.z > 0 && o.z > 0) return vof(1.0f / kPi);
  return vof(0.0f);
}
// locAt (Blinn exponent) (BSDF.hs:18-30) -- uses i.h where PBRT uses n.h; replicated as is.
inline Vec3 locAtBlinn(float exponent, Vec3 i, Vec3 o) {
  if (i.z > 0 && o.z > 0) {
    Vec3 h = norm(vadd(i, o));
    float cosThetaO = std::fabs(o.z);
    float cosThetaI = std::fabs(i.z);
    float cosThetaH = dot(i, h);
    float oDotH = dot(o, h);
    float d = ((exponent + 2.0f) * powf(std::fabs(h.z), exponent)) / (2.0f * kPi);
    float g = hs_min(1.0f, hs_min(((2.0f * cosThetaH) * cosThetaO) / oDotH,
                                  ((2.0f * cosThetaH) * cosThetaI) / oDotH));
    float f = 1.0f;                                         // fresnel ct = 1 -- TODO (BSDF.hs:34)
    return vof(((d * g) * f) / ((4.0f * cosThetaI) * cosThetaO));
  }
  return vof(0.0f);
}

// BSDF.at for the only BSDF the scene format can produce (Shaders.hs:12-14, main.hs:57-60):
//   Composite [Scaled diffuse Lambertian, Scaled specular (Blinn shininess)]
// evaluated as  sum [a, b] = (0 + a) + b  with Scaled = component-wise multiply (BSDF.hs:31-32).
inline Vec3 bsdfAt(const Material& m, const DifferentialGeometry& dg, Vec3 iw, Vec3 ow) {
  Vec3 nn = dg.dgNormal;                                    // BSDF.hs:36-38
  Vec3 sn = norm(dg.dgDPDU);
  Vec3 tn = cross(nn, sn);
  auto worldToLoc = [&](Vec3 v) { return Vec3{dot(v, sn), dot(v, tn), dot(v, nn)}; };  // :40
  Vec3 i = worldToLoc(iw), o = worldToLoc(ow);
  Vec3 a = vmul(m.diffuse, locAtLambertian(i, o));
  Vec3 b = vmul(m.specular, locAtBlinn(m.shininess, i, o));
  return vadd(vadd(vof(0.0f), a), b);
}

// ---------------------------------------------------------------------------------------
// EXTENSION (no reference counterpart, "parity unpinned"): spp > 1.  The reference shoots one
// ray through the integer raster coordinate (main.hs:73); sample 0 keeps offset (0,0) so that
// spp = 1 is exactly the reference.  Samples s >= 1 get a counter-based PCG-hash jitter in
// [0,1)^2.  The product kernel implements the same function.
// ---------------------------------------------------------------------------------------
inline uint32_t pcgHash(uint32_t x) {
  uint32_t state = x * 747796405u + 2891336453u;
  uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}
inline float sampleOffset(uint64_t seed, uint32_t pixel, uint32_t s, uint32_t dim) {
  if (s == 0) return 0.0f;
  uint32_t h = pcgHash((uint32_t)(seed >> 32) ^ 0x9E3779B9u);
  h = pcgHash(h ^ (uint32_t)seed);
  h = pcgHash(h ^ pixel);
  h = pcgHash(h ^ (s * 2u + dim));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------------------
// EXTENSION (no reference counterpart, "parity unpinned"): quad area lights.  The reference has
// point lights only (Lights.hs:7).  An area light is a one-sided parallelogram corner + a*edge1 +
// b*edge2 emitting `radiance` towards edge1 x edge2; per shading point it is sampled `samples`
// times with a counter-based PCG-hash (keyed by seed, pixel, pixel sample, recursion level, light
// slot), and each sample point acts as a point light of spectrum  cos_l * (area / samples) * radiance
// through the reference's own illuminationAtPoint / directIllumination arithmetic.  Light "slots"
// are numbered point lights first, then every sample of every area light.  The product kernels
// implement the same functions.
// ---------------------------------------------------------------------------------------
struct AreaLight {
  Vec3 corner, edge1, edge2, radiance;
  uint32_t samples;
  Vec3 normal;     // norm (edge1 x edge2)
  Vec3 flux;       // (area / samples) @* radiance
};
inline void finishAreaLight(AreaLight& a) {
  Vec3 n = cross(a.edge1, a.edge2);
  float area = len(n);
  a.normal = scale(1.0f / area, n);
  a.flux = scale(area / (float)a.samples, a.radiance);
}
struct ShadeCtx { uint64_t seed = 0; uint32_t pixel = 0, sample = 0, level = 0; };
inline float lightSampleU(const ShadeCtx& c, uint32_t slot, uint32_t dim) {
  uint32_t h = pcgHash((uint32_t)(c.seed >> 32) ^ 0x9E3779B9u);
  h = pcgHash(h ^ (uint32_t)c.seed);
  h = pcgHash(h ^ c.pixel);
  h = pcgHash(h ^ (0x80000000u | (c.sample * 2u + 1u)));
  h = pcgHash(h ^ ((c.level << 24) ^ (slot * 2u + dim)));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------------------
// Lights.hs:15-24, Integrators.hs
// ---------------------------------------------------------------------------------------
struct LightSample { Vec3 dir, intensity; Vec3 shadowOrigin, lightPos; };

inline Vec3 reflectionDir(Vec3 u, Vec3 n) {                 // Integrators.hs:46-47
  return vsub(u, scale(2.0f * dot(u, n), n));               // u - 2 * (u .* n) @* n
}

inline std::vector<AreaLight> areaLightsOf(const Scene& sc) {
  std::vector<AreaLight> out;
  for (size_t i = 0; i + 13 <= sc.areaLightData.size(); i += 13) {
    const float* f = sc.areaLightData.data() + i;
    AreaLight a;
    a.corner = Vec3{f[0], f[1], f[2]}; a.edge1 = Vec3{f[3], f[4], f[5]}; a.edge2 = Vec3{f[6], f[7], f[8]};
    a.radiance = Vec3{f[9], f[10], f[11]};
    a.samples = (uint32_t)f[12];
    finishAreaLight(a);
    out.push_back(a);
  }
  return out;
}

inline Vec3 directIllumination(const Scene& sc, const DifferentialGeometry& dg, const Ray& ray,
                               const Material& bsdf, Stats* st, const ShadeCtx& ctx = ShadeCtx()) {  // Integrators.hs:50-61
  Vec3 x = dg.dgPoint, n = dg.dgNormal;
  Vec3 total = vof(0.0f);                                   // sum = foldl (+) 0
  uint32_t slot = 0;
  for (const Light& light : sc.lights) {
    if (st) st->n_light++;
    // illuminationAtPoint (Lights.hs:15-24)
    Vec3 pointToLight = vsub(light.position, x);
    Vec3 lightDir = norm(pointToLight);
    Vec3 intensity = scale(1.0f / lensq(pointToLight), light.spectrum);
    Vec3 k = bsdfAt(bsdf, dg, lightDir, vneg(ray.u));
    Vec3 contrib = vof(0.0f);
    if (lensq(k) > 0) {                                     // && is lazy: probe only when needed
      if (st) st->n_shadow++;
      bool unoccluded = reachable(sc, vadd(x, scale(0.001f, lightDir)), light.position, st);
      if (unoccluded) contrib = vmul(scale(std::fabs(dot(lightDir, n)), k), intensity);
    }
    total = vadd(total, contrib);
    ++slot;
  }
  // extension: area lights, one virtual point light per sample
  if (!sc.areaLightData.empty()) {
    for (const AreaLight& al : areaLightsOf(sc)) {
      for (uint32_t j = 0; j < al.samples; ++j, ++slot) {
        if (st) st->n_light++;
        float u1 = lightSampleU(ctx, slot, 0), u2 = lightSampleU(ctx, slot, 1);
        Vec3 lightPos = vadd(vadd(al.corner, scale(u1, al.edge1)), scale(u2, al.edge2));
        Vec3 pointToLight = vsub(lightPos, x);
        Vec3 lightDir = norm(pointToLight);
        float cosL = -dot(lightDir, al.normal);
        Vec3 k = bsdfAt(bsdf, dg, lightDir, vneg(ray.u));
        Vec3 contrib = vof(0.0f);
        if (lensq(k) > 0 && cosL > 0) {
          if (st) st->n_shadow++;
          bool unoccluded = reachable(sc, vadd(x, scale(0.001f, lightDir)), lightPos, st);
          if (unoccluded) {
            Vec3 intensity = scale(1.0f / lensq(pointToLight), scale(cosL, al.flux));
            contrib = vmul(scale(std::fabs(dot(lightDir, n)), k), intensity);
          }
        }
        total = vadd(total, contrib);
      }
    }
  }
  return total;
}

// radiance / vcast / vhit (Integrators.hs:22-43).  `primOut`/`tOut` receive the primary hit.
inline Vec3 vcast(const Scene& sc, int maxDepth, const Ray& ray, Stats* st, int32_t* primOut = nullptr,
                  float* tOut = nullptr, ShadeCtx ctx = ShadeCtx()) {
  if (primOut) *primOut = -1;
  if (tOut) *tOut = 0.0f;
  if (maxDepth == 0) return vof(0.0f);                      // vcast 0 _ = Vec3 0 0 0
  std::optional<Hit> mh = rootCollider(sc, ray, st);
  if (!mh) return vof(0.0f);
  const Hit& hit = *mh;
  if (primOut) *primOut = hit.what;
  if (tOut) *tOut = hit.t;
  if (st) st->n_shade++;
  const Material& bsdf = sc.materials[sc.prims[hit.what].material];
  const DifferentialGeometry& dg = hit.dg;
  Vec3 x = dg.dgPoint, n = dg.dgNormal;
  Vec3 r = reflectionDir(ray.u, n);
  Vec3 fr = bsdfAt(bsdf, dg, r, vneg(ray.u));               // f r
  Ray next{vadd(x, scale(0.001f, r)), r, 1e6f};
  if (st && maxDepth - 1 > 0) st->n_secondary++;
  ShadeCtx deeper = ctx;
  deeper.level = ctx.level + 1;
  Vec3 rs = vcast(sc, maxDepth - 1, next, st, nullptr, nullptr, deeper);   // strict: always evaluated
  // (n .* r) @* f r * rs + directIllumination ...   ==  (((n.r) @* f r) * rs) + direct
  Vec3 refl = vmul(scale(dot(n, r), fr), rs);
  return vadd(refl, directIllumination(sc, dg, ray, bsdf, st, ctx));
}

// ---------------------------------------------------------------------------------------
// Cameras.hs (+ the `linear` package's !*!, !*, normalizePoint)
// ---------------------------------------------------------------------------------------
struct M44 { float m[4][4]; };

// linear: f !*! g = fmap (\f' -> foldl' (^+^) zero $ liftI2 (*^) f' g) f
inline M44 mmul(const M44& f, const M44& g) {
  M44 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float acc = 0.0f;
      for (int k = 0; k < 4; ++k) acc = acc + f.m[i][k] * g.m[k][j];
      r.m[i][j] = acc;
    }
  return r;
}
// linear: m !* v = fmap (\r -> sum $ liftI2 (*) r v) m ; sum = foldl' (+) 0
inline void mvmul(const M44& m, const float v[4], float out[4]) {
  for (int i = 0; i < 4; ++i) {
    float acc = 0.0f;
    for (int k = 0; k < 4; ++k) acc = acc + m.m[i][k] * v[k];
    out[i] = acc;
  }
}
inline M44 fromBasis(Vec3 a, Vec3 b, Vec3 c) {                                   // Cameras.hs:24-29
  return M44{{{a.x, b.x, c.x, 0}, {a.y, b.y, c.y, 0}, {a.z, b.z, c.z, 0}, {0, 0, 0, 1}}};
}
inline M44 translate(Vec3 t) {                                                   // Cameras.hs:32-37
  return M44{{{1, 0, 0, t.x}, {0, 1, 0, t.y}, {0, 0, 1, t.z}, {0, 0, 0, 1}}};
}
inline Vec3 transformPoint(const M44& tf, Vec3 p) {                              // Cameras.hs:40-44
  float v[4] = {p.x, p.y, p.z, 1.0f}, tv[4];
  mvmul(tf, v, tv);
  float iw = 1.0f / tv[3];                                  // normalizePoint (V4 a b c w) = (1/w) *^ V3 a b c
  return {iw * tv[0], iw * tv[1], iw * tv[2]};
}
struct Camera { float imW, imH, focalLength; Vec3 lookDir, upDir, position; };  // Cameras.hs:54-56

inline M44 cameraToWorld(Vec3 lookDir, Vec3 upDir, Vec3 position) {              // Cameras.hs:59-65
  Vec3 forward = norm(lookDir);
  Vec3 left = norm(cross(forward, upDir));
  Vec3 up = cross(left, forward);
  M44 dirTf = fromBasis(vneg(left), up, forward);
  return mmul(translate(position), dirTf);
}
inline M44 rasterToCamera(float imW, float imH) {                                // Cameras.hs:68-75
  float frameW = imW / imH, frameH = 1.0f;
  return M44{{{frameW / imW, 0, 0, -(frameW / 2.0f)},
              {0, -(frameH / imH), 0, frameH / 2.0f},
              {0, 0, 1, 0},
              {0, 0, 0, 1}}};
}
struct Caster { M44 tf, vtf; float focalLength; };
inline Caster makeCaster(const Camera& c) {                                      // Cameras.hs:78-82
  Caster k;
  k.tf = cameraToWorld(c.lookDir, c.upDir, c.position);
  k.vtf = mmul(k.tf, rasterToCamera(c.imW, c.imH));
  k.focalLength = c.focalLength;
  return k;
}
inline Ray computeInitialRay(const Caster& k, float u, float v) {                // Cameras.hs:83-86
  Vec3 origin = transformPoint(k.tf, Vec3{0, 0, 0});
  Vec3 direction = vsub(transformPoint(k.vtf, Vec3{u, v, k.focalLength}), origin);
  return Ray{origin, norm(direction), 1e6f};
}

// ---------------------------------------------------------------------------------------
// Sampling.hs, main.hs tiling
// ---------------------------------------------------------------------------------------
struct Window { int x0, y0, x1, y1; };

// batchWindow (Sampling.hs:9-21): the divisions default to Double.
inline Window batchWindow(int64_t w, int64_t h, int64_t num, int64_t count) {
  int64_t nx = count, ny = 1;
  while (nx % 2 == 0 && 2 * w * ny < h * nx) { nx = nx / 2; ny = ny * 2; }
  int64_t j = num / nx, i = num % nx;                       // quotRem (non-negative operands)
  Window r;
  r.x0 = (int)std::floor((double)i / (double)nx * (double)w);
  r.y0 = (int)std::floor((double)j / (double)ny * (double)h);
  r.x1 = (int)std::floor(((double)i + 1.0) / (double)nx * (double)w);
  r.y1 = (int)std::floor(((double)j + 1.0) / (double)ny * (double)h);
  return r;
}
// roundUpPow2 (main.hs:109-110): 2 ^ ceiling (logBase 2 x), logBase in Double = log x / log 2
inline int64_t roundUpPow2(int64_t x) {
  double e = std::ceil(std::log((double)x) / std::log(2.0));
  int64_t r = 1;
  for (int64_t i = 0; i < (int64_t)e; ++i) r *= 2;
  return r;
}
// nBatches (main.hs:128-130):  roundUpPow2 $ max (32 * numThreads) width * height `div` (16 * 16)
// parses as  ((max (32*numThreads) width) * height) `div` 256
inline int64_t numBatches(int64_t numThreads, int64_t width, int64_t height) {
  int64_t m = (32 * numThreads > width) ? 32 * numThreads : width;
  if ((m * height) / 256 < 1) return 1;   // the reference dies here (2 ^ negative); one tile instead
  return roundUpPow2((m * height) / 256);
}

}  // namespace yo
