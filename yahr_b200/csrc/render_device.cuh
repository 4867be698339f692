// render_device.cuh -- device functions of the per-ray path: camera ray, BVH traversal, ray-triangle
// and ray-sphere tests, BSDF, point lights, Whitted integrator.  Everything is IEEE binary32 with
// the reference's expression order; the translation unit is compiled with --fmad=false so no
// multiply-add is contracted, and divisions / square roots are the correctly rounded ones.
// Citations are to the reference's files.
#pragma once
#include "device_types.cuh"

namespace yb {
namespace dev {

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 xyz(const float4& a) { return mk(a.x, a.y, a.z); }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 vmul(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 vneg(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 vscale(float f, V3 v) { return mk(f * v.x, f * v.y, f * v.z); }      // (@*)
__device__ __forceinline__ float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }   // (.*)
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float rcp(float x) { return __fdiv_rn(1.0f, x); }
__device__ __forceinline__ float lensq(V3 v) { return dot(v, v); }
__device__ __forceinline__ float len(V3 v) { return __fsqrt_rn(dot(v, v)); }
__device__ __forceinline__ V3 vnormInline(V3 v) { return vscale(rcp(len(v)), v); }                   // norm
// CALL = true: the helper is a CALL to one out-of-line copy instead of an inline expansion -- same instructions, same
// bits.  The per-batch kernels (k_wf_fused*, k_wf_persist) use it for norm and BSDF.at: they keep warps of one SM in the
// walk, the shading and the probe walk at once, their hot code overflows the 32 KB L1.5 instruction cache, and the ten
// inline copies of norm and two of BSDF.at are what can be shared (C4 k_wf_fused 1.60 -> 1.47 ms, profiles/r2p).  The
// two-kernel set and the megakernel keep the inline expansions (outlined they are 4 - 7 % slower).
static __device__ __noinline__ V3 vnormCall(V3 v) { return vnormInline(v); }
template <bool CALL = false>
__device__ __forceinline__ V3 vnorm(V3 v) { return CALL ? vnormCall(v) : vnormInline(v); }
// GHC class-default min/max on Float (SURVEY.md note N)
__device__ __forceinline__ float hmin(float x, float y) { return x <= y ? x : y; }
__device__ __forceinline__ float hmax(float x, float y) { return x <= y ? y : x; }

struct Ray {
  V3 o, d;        // x0, u
  V3 inv;         // 1 / u per component (AABBs.hs:33 recomputes it per box; same value)
  bool exactNaN;  // some 1/u component is not finite: slab products can be NaN, use the select chain
};

__device__ __forceinline__ Ray makeRay(V3 o, V3 d) {
  Ray r;
  r.o = o; r.d = d;
  r.inv = mk(rcp(d.x), rcp(d.y), rcp(d.z));
  r.exactNaN = !(isfinite(r.inv.x) && isfinite(r.inv.y) && isfinite(r.inv.z));
  return r;
}

// bbRayIntersection (AABBs.hs:29-40) for the box (lo, hi) with the ray's current tMax.
// Returns `Just`-ness.  `key` is the value such that, for any later (smaller) tMax', the test
// would pass iff key <= tMax': the entry distance tNear, or -inf when a NaN slab product made
// the reference's tFar chain forget tMax.  This lets the test of a RIGHT child be evaluated
// when its parent is visited and re-validated when it is popped -- exactly equivalent to the
// reference's "test on entry with the current tMax".
__device__ __forceinline__ bool boxTest(float lox, float loy, float loz, float hix, float hiy, float hiz,
                                        const Ray& r, float tMax, float& key) {
  const float t0x = (lox - r.o.x) * r.inv.x, t1x = (hix - r.o.x) * r.inv.x;
  const float t0y = (loy - r.o.y) * r.inv.y, t1y = (hiy - r.o.y) * r.inv.y;
  const float t0z = (loz - r.o.z) * r.inv.z, t1z = (hiz - r.o.z) * r.inv.z;
  if (!r.exactNaN) {
    // no NaN can occur: the selects reduce to plain min / max (up to the sign of zero, which
    // no comparison below can observe)
    const float tN = fmaxf(fmaxf(fmaxf(0.0f, fminf(t0x, t1x)), fminf(t0y, t1y)), fminf(t0z, t1z));
    const float tF = fminf(fminf(fminf(tMax, fmaxf(t0x, t1x)), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    key = tN;
    return tN <= tF;
  }
  float tN = 0.0f, tF = tMax;
  tN = hmax(tN, hmin(t0x, t1x)); tF = hmin(tF, hmax(t0x, t1x));
  tN = hmax(tN, hmin(t0y, t1y)); tF = hmin(tF, hmax(t0y, t1y));
  tN = hmax(tN, hmin(t0z, t1z)); tF = hmin(tF, hmax(t0z, t1z));
  // hmax t0 t1 is NaN exactly when t0 is NaN; a NaN there replaces tFar and the next slab (or the
  // final comparison) no longer sees tMax.
  const bool forgets = (t0x != t0x) || (t0y != t0y) || (t0z != t0z);
  key = forgets ? -INFINITY : tN;
  return !(tN > tF);
}

// collideTriangle / collideSphere acceptance (Shapes.hs:13-29, 36-59) for the primitive at DFS
// position idx.  Returns true and t iff the reference returns `Just`.
__device__ __forceinline__ bool hitPrimitive(const DeviceScene& sc, uint32_t idx, const Ray& r, float tMax,
                                             float& tOut, uint32_t* candidates = nullptr) {
  const float4 A = __ldg(&sc.prims[3 * (size_t)idx + 0]);
  const float4 B = __ldg(&sc.prims[3 * (size_t)idx + 1]);
  if (__float_as_uint(A.w) & 1u) {
    const float4 C = __ldg(&sc.prims[3 * (size_t)idx + 2]);
    const V3 p0 = xyz(A), e1 = xyz(B), e2 = xyz(C);
    const V3 s = vsub(r.o, p0);
    const V3 s1 = cross(r.d, e2);
    const V3 s2 = cross(s, e1);
    const float invDiv = rcp(dot(s1, e1));
    const float t = dot(s2, e2) * invDiv;
    const float b1 = dot(s1, s) * invDiv;
    const float b2 = dot(s2, r.d) * invDiv;
    const float b0 = (1.0f - b1) - b2;
    if (!(b0 >= 0.0f && b0 <= 1.0f && b1 >= 0.0f && b1 <= 1.0f && b2 >= 0.0f && b2 <= 1.0f && t > 0.0f &&
          t <= tMax))
      return false;
    if (candidates) ++*candidates;                    // counting build: the normals are fetched for this test
    const V3 n0 = xyz(__ldg(&sc.normals[3 * (size_t)idx + 0]));
    const V3 n1 = xyz(__ldg(&sc.normals[3 * (size_t)idx + 1]));
    const V3 n2 = xyz(__ldg(&sc.normals[3 * (size_t)idx + 2]));
    const V3 ns = vadd(vadd(vscale(b0, n0), vscale(b1, n1)), vscale(b2, n2));
    if (!(dot(r.d, ns) < 0.0f)) return false;
    tOut = t;
    return true;
  }
  const V3 c = xyz(A);
  const float rad = B.x;
  const V3 d = vsub(r.o, c);
  const float a = dot(r.d, r.d);
  const float b = dot(vmul(mk(2.0f, 2.0f, 2.0f), d), r.d);
  const float cc = dot(d, d) - rad * rad;
  const float delta = b * b - (4.0f * a) * cc;
  if (delta < 0.0f) return false;
  const float sq = __fsqrt_rn(delta);
  const float den = 2.0f * a;
  const float ta = __fdiv_rn(-b - sq, den);
  if (ta > 0.0f && ta <= tMax) { tOut = ta; return true; }
  const float tb = __fdiv_rn(-b + sq, den);
  if (tb > 0.0f && tb <= tMax) { tOut = tb; return true; }
  return false;
}

static const uint32_t kNoHit = 0xFFFFFFFFu;

// The root collider (Culling.hs:24-25, 33, 38, 52) on the flattened tree.
//   ANY_HIT : stop at the first accepted primitive (`reachable` only uses emptiness, Rays.hs:50;
//             up to the first acceptance the walk is identical to the closest-hit one).
//   ORDERED : visit the nearer child first; equal-t ties go to the larger DFS position, which is
//             what "every later Just replaces the hit" (Rays.hs:43-46) produces in the reference.
// Returns the DFS position of the hit (kNoHit for Nothing) and its t.
template <bool ANY_HIT, bool ORDERED>
__device__ __forceinline__ uint32_t traverse(const DeviceScene& sc, const Ray& r, float tMax, float& tHit) {
  uint32_t best = kNoHit;
  if (sc.rootRef == kDevRefNull) return best;
  float key;
  if (!boxTest(sc.rootLo[0], sc.rootLo[1], sc.rootLo[2], sc.rootHi[0], sc.rootHi[1], sc.rootHi[2], r, tMax, key))
    return best;

  uint2 stack[64];
  int sp = 0;
  uint32_t cur = sc.rootRef;
  for (;;) {
    if (!(cur & kDevRefLeafBit)) {
      const float4* np = sc.nodes + 4 * (size_t)cur;
      const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
      float keyL, keyR;
      const uint32_t refL = __float_as_uint(n3.x), refR = __float_as_uint(n3.y);
      const bool passL = boxTest(n0.x, n0.y, n2.x, n0.z, n0.w, n2.y, r, tMax, keyL) && refL != kDevRefNull;
      const bool passR = boxTest(n1.x, n1.y, n2.z, n1.z, n1.w, n2.w, r, tMax, keyR) && refR != kDevRefNull;
      if (passL && passR) {
        if (ORDERED && keyR < keyL) {
          stack[sp++] = make_uint2(refL, __float_as_uint(keyL));
          cur = refR;
        } else {
          stack[sp++] = make_uint2(refR, __float_as_uint(keyR));
          cur = refL;
        }
        continue;
      }
      if (passL) { cur = refL; continue; }
      if (passR) { cur = refR; continue; }
    } else {
      uint32_t first, count;
      if ((cur & kDevRefMultiBits) == kDevRefMultiBits) {
        const uint2 ml = __ldg(&sc.multiLeaves[cur & 0x3FFFFFFFu]);
        first = ml.x; count = ml.y;
      } else {
        first = cur & 0x3FFFFFFFu; count = 1;
      }
      for (uint32_t k = 0; k < count; ++k) {           // collideAll over the leaf's primitives, in order
        const uint32_t idx = first + k;
        float t;
        if (hitPrimitive(sc, idx, r, tMax, t)) {
          // reference order: every Just replaces (later DFS position wins equal t).  In ORDERED
          // mode positions are not visited monotonically, so the tie is decided explicitly.
          if (!ORDERED || t < tMax || best == kNoHit || idx > best) {
            best = idx; tMax = t;
            if (ANY_HIT) { tHit = t; return best; }
          }
        }
      }
    }
    // pop: re-validate the deferred box test against the current tMax
    for (;;) {
      if (sp == 0) { tHit = tMax; return best; }
      const uint2 e = stack[--sp];
      if (__uint_as_float(e.y) <= tMax) { cur = e.x; break; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Shading (BSDF.hs, Shaders.hs, Lights.hs, Integrators.hs)
// ---------------------------------------------------------------------------------------------
struct Surface {
  V3 x, n, dpdu;       // dgPoint, dgNormal, dgDPDU (dgDPDV is never read downstream)
  uint32_t material, primId;
};

// DifferentialGeometry of the accepted hit: recomputes the winner's intermediate values with the
// same instructions as hitPrimitive, hence the same bits (Shapes.hs:20-27, 46-55).
template <bool CALL = false>
__device__ __forceinline__ Surface surfaceAt(const DeviceScene& sc, uint32_t idx, const Ray& r, float t) {
  Surface s;
  const float4 A = __ldg(&sc.prims[3 * (size_t)idx + 0]);
  const float4 B = __ldg(&sc.prims[3 * (size_t)idx + 1]);
  const uint32_t meta = __float_as_uint(A.w);
  s.material = meta >> 1;
  s.primId = __float_as_uint(B.w);
  s.x = vadd(r.o, vscale(t, r.d));                     // rayAt / x0 + t @* u
  if (meta & 1u) {
    const float4 C = __ldg(&sc.prims[3 * (size_t)idx + 2]);
    const V3 p0 = xyz(A), e1 = xyz(B), e2 = xyz(C);
    const V3 sv = vsub(r.o, p0);
    const V3 s1 = cross(r.d, e2);
    const V3 s2 = cross(sv, e1);
    const float invDiv = rcp(dot(s1, e1));
    const float b1 = dot(s1, sv) * invDiv;
    const float b2 = dot(s2, r.d) * invDiv;
    const float b0 = (1.0f - b1) - b2;
    const V3 n0 = xyz(__ldg(&sc.normals[3 * (size_t)idx + 0]));
    const V3 n1 = xyz(__ldg(&sc.normals[3 * (size_t)idx + 1]));
    const V3 n2 = xyz(__ldg(&sc.normals[3 * (size_t)idx + 2]));
    const V3 ns = vadd(vadd(vscale(b0, n0), vscale(b1, n1)), vscale(b2, n2));
    const V3 ss = vnorm<CALL>(e2);
    const V3 ts = vnorm<CALL>(cross(ss, ns));
    s.n = ns;
    s.dpdu = cross(ts, ns);
  } else {
    const V3 c = xyz(A);
    const V3 n = vnorm<CALL>(vsub(s.x, c));
    s.n = n;
    s.dpdu = cross(n, mk(0.0f, 0.0f, 1.0f));
  }
  return s;
}

struct MaterialD { V3 diffuse, specular; float shininess; };

__device__ __forceinline__ MaterialD loadMaterial(const DeviceScene& sc, uint32_t m) {
  const float4 a = __ldg(&sc.materials[2 * m + 0]);
  const float4 b = __ldg(&sc.materials[2 * m + 1]);
  MaterialD r;
  r.diffuse = xyz(a); r.shininess = a.w; r.specular = xyz(b);
  return r;
}

#define YB_PI 3.14159274101257324f   /* pi :: Float */

// Shading frame (BSDF.hs:36-40)
struct Frame { V3 sn, tn, nn; };
template <bool CALL = false>
__device__ __forceinline__ Frame makeFrame(const Surface& s) {
  Frame f;
  f.nn = s.n;
  f.sn = vnorm<CALL>(s.dpdu);
  f.tn = cross(f.nn, f.sn);
  return f;
}
__device__ __forceinline__ V3 toLocal(const Frame& f, V3 v) { return mk(dot(v, f.sn), dot(v, f.tn), dot(v, f.nn)); }

// BSDF.at for Composite [Scaled diffuse Lambertian, Scaled specular (Blinn e)] (Shaders.hs:12-14,
// BSDF.hs:12-46): (0 + diffuse * lambert) + specular * blinn, all in the local frame.
template <bool CALL>
__device__ __forceinline__ V3 bsdfAtBody(const MaterialD& m, const Frame& f, V3 iw, V3 ow) {
  const V3 i = toLocal(f, iw), o = toLocal(f, ow);
  float lam = 0.0f, bl = 0.0f;
  if (i.z > 0.0f && o.z > 0.0f) {
    lam = __fdiv_rn(1.0f, YB_PI);
    const V3 h = vnorm<CALL>(vadd(i, o));
    const float cosThetaO = fabsf(o.z), cosThetaI = fabsf(i.z);
    const float cosThetaH = dot(i, h);
    const float oDotH = dot(o, h);
    const float d = __fdiv_rn((m.shininess + 2.0f) * powf(fabsf(h.z), m.shininess), 2.0f * YB_PI);
    const float g = hmin(1.0f, hmin(__fdiv_rn((2.0f * cosThetaH) * cosThetaO, oDotH),
                                    __fdiv_rn((2.0f * cosThetaH) * cosThetaI, oDotH)));
    bl = __fdiv_rn((d * g) * 1.0f, (4.0f * cosThetaI) * cosThetaO);
  }
  const V3 a = vmul(m.diffuse, mk(lam, lam, lam));
  const V3 b = vmul(m.specular, mk(bl, bl, bl));
  return vadd(vadd(mk(0.0f, 0.0f, 0.0f), a), b);
}
static __device__ __noinline__ V3 bsdfAtCall(const MaterialD& m, const Frame& f, V3 iw, V3 ow) { return bsdfAtBody<true>(m, f, iw, ow); }
template <bool CALL = false>
__device__ __forceinline__ V3 bsdfAt(const MaterialD& m, const Frame& f, V3 iw, V3 ow) {
  return CALL ? bsdfAtCall(m, f, iw, ow) : bsdfAtBody<false>(m, f, iw, ow);
}

// Counter-based jitter for samples >= 1 (extension; sample 0 = the reference's ray).
__device__ __forceinline__ uint32_t pcgHash(uint32_t x) {
  const uint32_t state = x * 747796405u + 2891336453u;
  const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}
__device__ __forceinline__ float sampleOffset(uint64_t seed, uint32_t pixel, uint32_t s, uint32_t dim) {
  if (s == 0) return 0.0f;
  uint32_t h = pcgHash((uint32_t)(seed >> 32) ^ 0x9E3779B9u);
  h = pcgHash(h ^ (uint32_t)seed);
  h = pcgHash(h ^ pixel);
  h = pcgHash(h ^ (s * 2u + dim));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// Extension: quad area lights (include/yahr_b200.h).  The coordinates
// of the sample point of light slot `slot` on its parallelogram, keyed by pixel / pixel sample / level.
struct ShadeCtx { uint64_t seed; uint32_t pixel, sample, level; };
__device__ __forceinline__ float lightSampleU(const ShadeCtx& c, uint32_t slot, uint32_t dim) {
  uint32_t h = pcgHash((uint32_t)(c.seed >> 32) ^ 0x9E3779B9u);
  h = pcgHash(h ^ (uint32_t)c.seed);
  h = pcgHash(h ^ c.pixel);
  h = pcgHash(h ^ (0x80000000u | (c.sample * 2u + 1u)));
  h = pcgHash(h ^ ((c.level << 24) ^ (slot * 2u + dim)));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}
struct AreaLightD { V3 corner, edge1, edge2, normal, flux; uint32_t samples; };
__device__ __forceinline__ AreaLightD loadAreaLight(const DeviceScene& sc, uint32_t a) {
  const float4 r0 = __ldg(&sc.areaLights[5 * a + 0]);
  AreaLightD l;
  l.corner = xyz(r0); l.samples = __float_as_uint(r0.w);
  l.edge1 = xyz(__ldg(&sc.areaLights[5 * a + 1])); l.edge2 = xyz(__ldg(&sc.areaLights[5 * a + 2]));
  l.normal = xyz(__ldg(&sc.areaLights[5 * a + 3])); l.flux = xyz(__ldg(&sc.areaLights[5 * a + 4]));
  return l;
}
__device__ __forceinline__ V3 areaLightPoint(const AreaLightD& l, const ShadeCtx& ctx, uint32_t slot) {
  const float u1 = lightSampleU(ctx, slot, 0), u2 = lightSampleU(ctx, slot, 1);
  return vadd(vadd(l.corner, vscale(u1, l.edge1)), vscale(u2, l.edge2));
}

struct Counters { uint32_t primary, shadow, secondary; };

// directIllumination (Integrators.hs:50-61) + illuminationAtPoint (Lights.hs:15-24) + reachable (Rays.hs:49-54)
// for one light slot: a point light (spectrum = its spectrum) or one sample of an area light (extension:
// spectrum = its flux, scaled by the emitter cosine; nothing is emitted from the back side).
template <bool ORDERED>
__device__ __forceinline__ V3 slotIllumination(const DeviceScene& sc, const Surface& s, const Frame& f,
                                               const MaterialD& m, V3 wo, V3 lightPos, V3 spectrum, bool area,
                                               V3 lightNormal, Counters& cnt) {
  const V3 pointToLight = vsub(lightPos, s.x);
  const V3 lightDir = vnorm(pointToLight);
  const V3 k = bsdfAt(m, f, lightDir, wo);
  const float cosL = area ? -dot(lightDir, lightNormal) : 1.0f;
  V3 contrib = mk(0.0f, 0.0f, 0.0f);
  if (lensq(k) > 0.0f && (!area || cosL > 0.0f)) {
    const V3 p0 = vadd(s.x, vscale(0.001f, lightDir));
    const V3 dl = vsub(lightPos, p0);
    const Ray probe = makeRay(p0, vnorm(dl));
    float th;
    cnt.shadow++;
    const bool occluded = traverse<true, ORDERED>(sc, probe, len(dl), th) != kNoHit;
    if (!occluded) {
      const V3 intensity = vscale(rcp(lensq(pointToLight)), area ? vscale(cosL, spectrum) : spectrum);
      contrib = vmul(vscale(fabsf(dot(lightDir, s.n)), k), intensity);
    }
  }
  return contrib;
}

template <bool ORDERED>
__device__ __forceinline__ V3 directIllumination(const DeviceScene& sc, const Surface& s, const Frame& f,
                                                 const MaterialD& m, V3 rayDir, Counters& cnt, const ShadeCtx& ctx) {
  V3 total = mk(0.0f, 0.0f, 0.0f);
  const V3 wo = vneg(rayDir);
  uint32_t slot = 0;
  for (uint32_t li = 0; li < sc.nLights; ++li, ++slot) {
    const V3 lightPos = xyz(__ldg(&sc.lights[2 * li + 0]));
    const V3 spectrum = xyz(__ldg(&sc.lights[2 * li + 1]));
    total = vadd(total, slotIllumination<ORDERED>(sc, s, f, m, wo, lightPos, spectrum, false, mk(0, 0, 0), cnt));
  }
  for (uint32_t a = 0; a < sc.nAreaLights; ++a) {
    const AreaLightD l = loadAreaLight(sc, a);
    for (uint32_t j = 0; j < l.samples; ++j, ++slot)
      total = vadd(total, slotIllumination<ORDERED>(sc, s, f, m, wo, areaLightPoint(l, ctx, slot), l.flux, true,
                                                    l.normal, cnt));
  }
  return total;
}

// computeInitialRay (Cameras.hs:83-86) with the `linear` summation order ((0 + a) + b) + c) + d.
template <bool CALL = false>
__device__ __forceinline__ Ray cameraRay(const RenderParams& P, float u, float v) {
  float p[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = 0.0f + P.vtf[4 * i + 0] * u;
    acc = acc + P.vtf[4 * i + 1] * v;
    acc = acc + P.vtf[4 * i + 2] * P.focal;
    acc = acc + P.vtf[4 * i + 3] * 1.0f;
    p[i] = acc;                                   // normalizePoint: w == 1 exactly (checked on the host)
  }
  const V3 origin = mk(P.origin[0], P.origin[1], P.origin[2]);
  const V3 direction = vsub(mk(p[0], p[1], p[2]), origin);
  return makeRay(origin, vnorm<CALL>(direction));
}

// radiance / vcast / vhit (Integrators.hs:22-43), recursion unrolled into a forward pass that
// records (weight, direct) per level and a backward fold  acc = weight * acc + direct, which is
// the reference's  ((n.r) @* f r) * rs + direct  evaluated innermost first.
template <bool ORDERED>
__device__ __forceinline__ V3 radiance(const RenderParams& P, Ray ray, Counters& cnt, uint32_t& primOut, ShadeCtx ctx) {
  V3 weight[16], direct[16];
  int levels = 0;
  primOut = kNoHit;
  for (int level = 0; level < P.depth; ++level) {
    float t;
    const uint32_t idx = traverse<false, ORDERED>(P.sc, ray, 1e6f, t);
    if (idx == kNoHit) break;                                 // Nothing -> Vec3 0 0 0
    const Surface s = surfaceAt(P.sc, idx, ray, t);
    if (level == 0) primOut = s.primId;
    const MaterialD m = loadMaterial(P.sc, s.material);
    const Frame f = makeFrame(s);
    const V3 refl = vsub(ray.d, vscale(2.0f * dot(ray.d, s.n), s.n));     // reflectionDir
    const V3 fr = bsdfAt(m, f, refl, vneg(ray.d));
    weight[levels] = vscale(dot(s.n, refl), fr);
    ctx.level = (uint32_t)level;
    direct[levels] = directIllumination<ORDERED>(P.sc, s, f, m, ray.d, cnt, ctx);
    ++levels;
    if (level + 1 < P.depth) {
      cnt.secondary++;
      ray = makeRay(vadd(s.x, vscale(0.001f, refl)), refl);
    }
  }
  V3 acc = mk(0.0f, 0.0f, 0.0f);
  for (int k = levels - 1; k >= 0; --k) acc = vadd(vmul(weight[k], acc), direct[k]);
  return acc;
}

}  // namespace dev
}  // namespace yb
