// wavefront.cu -- persistent-thread wavefront kernel set (recursion depth 1 = direct lighting, the
// mode of every BASELINE config).  Compile with --fmad=false.
//
//   k_wf_primary : persistent warps pull batches of 32 work items (pixels, tile-major order, 8x4 pixel
//                  footprint per warp) from a global counter, generate the camera rays (A3), run the
//                  closest-hit walk (A4-A8) and -- once the whole warp has converged -- shade the hits
//                  (A9-A13): DifferentialGeometry, material, BSDF per light; the pixel's base value is
//                  stored and one shadow-probe record per light with a non-zero BSDF
//                  (Integrators.hs:59) is appended to a queue (ballot/popc compaction, one atomic per warp).
//   k_wf_shadow  : persistent warps run any-hit walks over the compacted queue; an unoccluded probe
//                  writes its contribution to the pixel (single light) or sets its visibility flag
//                  (several lights; k_wf_resolve then sums a pixel's lights in light order).
//   k_wf_accum   : spp > 1 only (extension): one launch traces up to 16 samples of every pixel (sample-major
//                  work items, per-sample frames); acc += sample in sample order, and the final divide.
//   k_wf_fused   : ONE light slot, per batch: the warp that traced and shaded a batch walks its probes itself; the
//                  default for frames / shares of <= 2.5 M work items and behind the streamed float host rows;
//                  k_wf_fused_depth{,_lights} carry the same plan through the levels of Whitted recursion.
//   Selectable, parity-tested, not default ("Round 2"): k_wf_persist (one launch per frame, probes
//   through committed chunks in per-CTA rings), k_wf_trace + k_wf_shade (walk and shading as two kernels), the walk
//   on compressed 64-byte nodes (template flag CMP), the first stack entries in shared memory (Stack<SH>).
//
// The walk is the reference's (left child first, box test on entry with the current tMax, later hit
// replaces), by default on the 4-wide collapse of the tree (wide_bvh.cu: same leaves, same order, same
// tMax at every leaf; rays with a non-finite 1/u stay on the binary tree), with an inner phase specialised
// per ray octant.  Rays visit ~10 wide nodes per leaf, so at any moment only a few lanes of a warp hold a
// leaf; running the (long) primitive-test code for them every round wastes most of the warp.  Lanes
// that reach a leaf therefore PARK until at least `leafThreshold` lanes hold one (or no lane has inner
// work left); then the leaf code runs once for all of them.  Parking only delays a lane -- each ray
// still executes its own steps in the reference order, so results are unchanged.
#include "kernels.hpp"
#include "render_device.cuh"

// This file is compiled twice: as it stands (the production kernels), and through wavefront_count.cu with
// YB_COUNT_WORK defined -- the COUNTING build: the same kernels in namespace yb::counted with per-lane work counters
// (visited nodes, primitive tests, normal fetches, stack pushes / pops, shaded hits, probes) summed into
// WavefrontParams::workStats.  The production kernels contain no trace of the counters.
#ifdef YB_COUNT_WORK
#define YB_WF_OPEN namespace yb { namespace counted {
#define YB_WF_CLOSE } }
#define YB_CNT(s, k) (++(s).cnt[k])
#define YB_CNT_INIT(s) do { for (int k_ = 0; k_ < kNumCnt; ++k_) (s).cnt[k_] = 0u; } while (0)
#else
#define YB_WF_OPEN namespace yb {
#define YB_WF_CLOSE }
#define YB_CNT(s, k) ((void)0)
#define YB_CNT_INIT(s) ((void)0)
#endif

YB_WF_OPEN

using namespace dev;

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kStackEntries = 64;

// per-lane work counters of the counting build (indices into Trav::cnt and, per walk kind, WavefrontParams::workStats)
enum { kCntWide = 0, kCntBinary, kCntPrim, kCntCand, kCntPush, kCntPop, kNumCnt };

struct Trav {
  uint32_t cur;
  int sp;
  float tMax;
  uint32_t best;
#ifdef YB_COUNT_WORK
  uint32_t cnt[kNumCnt];
#endif
};

// Traversal stack of a lane: (subtree reference, entry distance) pairs.  The first SH entries live in SHARED memory,
// entry i of thread t at column t of row i (row = 128 x 8 B): whatever the lanes' stack pointers are, lane l always
// hits banks 2l, 2l+1 -- every push / pop is conflict-free (two 128 B wavefronts per warp) and stays out of the L1
// tag stage and of the L1 lines the tree lives in.  Deeper entries (rare) spill to local memory.  SH = 0: the whole
// stack in local memory (round 1: 35 % of k_wf_primary's L1 sectors were stack traffic, profiles/r1r).
template <int SH>
struct Stack {
  uint2* sh;
  uint2* lo;
  __device__ __forceinline__ void store(int i, uint2 e) const {
    if (SH == 0) lo[i] = e;
    else if (i < SH) sh[i * 128] = e;
    else lo[i - SH] = e;
  }
  __device__ __forceinline__ uint2 load(int i) const {
    if (SH == 0) return lo[i];
    if (i < SH) return sh[i * 128];
    return lo[i - SH];
  }
};
#define YB_STACK(SH)                                                     \
  __shared__ uint2 shStack_[(SH) > 0 ? (SH) * 128 : 1];                  \
  uint2 loStack_[kStackEntries - (SH)];                                  \
  const Stack<(SH)> stack{shStack_ + threadIdx.x, loStack_}

// Pops the next deferred subtree whose box test still passes with the current tMax.
template <class STK>
__device__ __forceinline__ bool popNext(Trav& s, const STK& stack) {
  while (s.sp > 0) {
    const uint2 e = stack.load(--s.sp);
    YB_CNT(s, kCntPop);
    if (__uint_as_float(e.y) <= s.tMax) { s.cur = e.x; return true; }
  }
  return false;
}

// Starts a walk: root box test (wrapCollider of the root, AABBs.hs:42-43).  Returns false if the
// ray is finished immediately.
__device__ __forceinline__ bool travBegin(const DeviceScene& sc, const Ray& r, float tMax, Trav& s) {
  s.sp = 0; s.tMax = tMax; s.best = kNoHit; s.cur = sc.rootRef;
  if (sc.rootRef == kDevRefNull) return false;
  float key;
  return boxTest(sc.rootLo[0], sc.rootLo[1], sc.rootLo[2], sc.rootHi[0], sc.rootHi[1], sc.rootHi[2], r, tMax, key);
}

// Ray constants of the packed slab test: (-o) and 1/u as register pairs for FADD2 / FMUL2.
struct RayPack {
  float2 negOxy, negOzz, invXy, invZz;
};
__device__ __forceinline__ RayPack packRay(const Ray& r) {
  RayPack p;
  p.negOxy = make_float2(-r.o.x, -r.o.y);
  p.negOzz = make_float2(-r.o.z, -r.o.z);
  p.invXy = make_float2(r.inv.x, r.inv.y);
  p.invZz = make_float2(r.inv.z, r.inv.z);
  return p;
}

// Octant of a ray whose 1/u components are all finite: bit k set <=> 1/u_k < 0.
__device__ __forceinline__ int rayOctant(const Ray& r) {
  return (r.inv.x < 0.0f ? 1 : 0) | (r.inv.y < 0.0f ? 2 : 0) | (r.inv.z < 0.0f ? 4 : 0);
}

// Both child box tests of one node for a ray of octant OCT (all 1/u components finite, so no NaN
// can occur and key = tNear).  Same arithmetic as bbRayIntersection (AABBs.hs:29-40):
//   t = (b - o) * (1/u)  as  (b + (-o)) * inv  -- identical in IEEE, done as packed f32x2 ops;
//   min t0 t1 / max t0 t1: for 1/u > 0 the subtraction and the multiplication are monotone, so
//   t0 <= t1 and the min IS t0 (1/u < 0: t1).  The octant picks them at compile time.
template <int OCT>
__device__ __forceinline__ void childTestsOct(const float4& n0, const float4& n1, const float4& n2, const RayPack& p,
                                              float tMax, bool& passL, float& keyL, bool& passR, float& keyR) {
  const float2 l0 = __fmul2_rn(__fadd2_rn(make_float2(n0.x, n0.y), p.negOxy), p.invXy);   // L: t0x t0y
  const float2 l1 = __fmul2_rn(__fadd2_rn(make_float2(n0.z, n0.w), p.negOxy), p.invXy);   // L: t1x t1y
  const float2 r0 = __fmul2_rn(__fadd2_rn(make_float2(n1.x, n1.y), p.negOxy), p.invXy);   // R: t0x t0y
  const float2 r1 = __fmul2_rn(__fadd2_rn(make_float2(n1.z, n1.w), p.negOxy), p.invXy);   // R: t1x t1y
  const float2 lz = __fmul2_rn(__fadd2_rn(make_float2(n2.x, n2.y), p.negOzz), p.invZz);   // L: t0z t1z
  const float2 rz = __fmul2_rn(__fadd2_rn(make_float2(n2.z, n2.w), p.negOzz), p.invZz);   // R: t0z t1z
  constexpr bool NX = (OCT & 1) != 0, NY = (OCT & 2) != 0, NZ = (OCT & 4) != 0;
  const float tNL = fmaxf(fmaxf(fmaxf(0.0f, NX ? l1.x : l0.x), NY ? l1.y : l0.y), NZ ? lz.y : lz.x);
  const float tFL = fminf(fminf(fminf(tMax, NX ? l0.x : l1.x), NY ? l0.y : l1.y), NZ ? lz.x : lz.y);
  const float tNR = fmaxf(fmaxf(fmaxf(0.0f, NX ? r1.x : r0.x), NY ? r1.y : r0.y), NZ ? rz.y : rz.x);
  const float tFR = fminf(fminf(fminf(tMax, NX ? r0.x : r1.x), NY ? r0.y : r1.y), NZ ? rz.x : rz.y);
  keyL = tNL; passL = tNL <= tFL;
  keyR = tNR; passR = tNR <= tFR;
}

// One inner-node visit.  OCT in 0..7: every busy lane of the warp has that octant (packed fast path);
// OCT = -1: generic path (mixed octants, or some ray whose slab products can be NaN).
// Returns false when the walk is finished (nothing left to visit).
template <bool ORDERED, int OCT, class STK>
__device__ __forceinline__ bool innerStep(const DeviceScene& sc, const Ray& r, const RayPack& rp, Trav& s,
                                          const STK& stack) {
  YB_CNT(s, kCntBinary);
  const float4* np = sc.nodes + 4 * (size_t)s.cur;
  const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
  float keyL, keyR;
  bool passL, passR;
  const uint32_t refL = __float_as_uint(n3.x), refR = __float_as_uint(n3.y);
  if (OCT >= 0) {
    childTestsOct<(OCT >= 0 ? OCT : 0)>(n0, n1, n2, rp, s.tMax, passL, keyL, passR, keyR);
  } else {
    passL = boxTest(n0.x, n0.y, n2.x, n0.z, n0.w, n2.y, r, s.tMax, keyL);
    passR = boxTest(n1.x, n1.y, n2.z, n1.z, n1.w, n2.w, r, s.tMax, keyR);
  }
  passL = passL && refL != kDevRefNull;
  passR = passR && refR != kDevRefNull;
  if (passL && passR) {
    YB_CNT(s, kCntPush);
    if (ORDERED && keyR < keyL) {
      stack.store(s.sp++, make_uint2(refL, __float_as_uint(keyL)));
      s.cur = refR;
    } else {
      stack.store(s.sp++, make_uint2(refR, __float_as_uint(keyR)));
      s.cur = refL;
    }
    return true;
  }
  if (passL) { s.cur = refL; return true; }
  if (passR) { s.cur = refR; return true; }
  return popNext(s, stack);
}

// One leaf visit (collideAll over the leaf's primitives in order).  Returns false when finished.
// anyRt: any-hit decided at run time (k_wf_fused shares ONE copy of the walk between its closest-hit and any-hit phases).
template <bool ANY_HIT, bool ORDERED, class STK>
__device__ __forceinline__ bool leafStep(const DeviceScene& sc, const Ray& r, Trav& s, const STK& stack, bool anyRt = false) {
  uint32_t first, count;
  if ((s.cur & kDevRefMultiBits) == kDevRefMultiBits) {
    const uint2 ml = __ldg(&sc.multiLeaves[s.cur & 0x3FFFFFFFu]);
    first = ml.x; count = ml.y;
  } else {
    first = s.cur & 0x3FFFFFFFu; count = 1;
  }
  for (uint32_t k = 0; k < count; ++k) {
    const uint32_t idx = first + k;
    float t;
    YB_CNT(s, kCntPrim);
#ifdef YB_COUNT_WORK
    if (hitPrimitive(sc, idx, r, s.tMax, t, &s.cnt[kCntCand])) {
#else
    if (hitPrimitive(sc, idx, r, s.tMax, t)) {
#endif
      if (!ORDERED || t < s.tMax || s.best == kNoHit || idx > s.best) {
        s.best = idx; s.tMax = t;
        if (ANY_HIT || anyRt) return false;
      }
    }
  }
  return popNext(s, stack);
}

// Runs the walks of a whole warp to completion with leaf parking.  `busy` per lane.
template <bool ANY_HIT, bool ORDERED, int OCT, class STK>
__device__ __forceinline__ void traverseWarpOct(const DeviceScene& sc, const Ray& r, Trav& s, const STK& stack, bool busy,
                                             int leafThreshold, bool anyRt = false) {
  const RayPack rp = packRay(r);
  for (;;) {
    const bool atLeaf = busy && (s.cur & kDevRefLeafBit);
    const bool atInner = busy && !atLeaf;
    const unsigned mLeaf = __ballot_sync(kFull, atLeaf);
    const unsigned mInner = __ballot_sync(kFull, atInner);
    if ((mLeaf | mInner) == 0) break;
    if (mInner == 0 || __popc(mLeaf) >= leafThreshold) {
      if (atLeaf) busy = leafStep<ANY_HIT, ORDERED>(sc, r, s, stack, anyRt);
    }
    if (atInner) busy = innerStep<ORDERED, OCT>(sc, r, rp, s, stack);
  }
}

// Picks the octant-specialised walk when every busy lane of the warp shares one octant and no lane
// can produce NaN slab products; otherwise the generic walk.
template <bool ANY_HIT, bool ORDERED, class STK>
__device__ __forceinline__ void traverseWarp(const DeviceScene& sc, const Ray& r, Trav& s, const STK& stack, bool busy,
                                             int leafThreshold, bool packed) {
  const unsigned mBusy = __ballot_sync(kFull, busy);
  if (mBusy == 0) return;
  int oct = -1;
  if (packed) {
    const int mine = rayOctant(r);
    const int lead = __shfl_sync(kFull, mine, __ffs(mBusy) - 1);
    if (__all_sync(kFull, !busy || (mine == lead && !r.exactNaN))) oct = lead;
  }
  switch (oct) {
    case 0: traverseWarpOct<ANY_HIT, ORDERED, 0>(sc, r, s, stack, busy, leafThreshold); break;
    case 1: traverseWarpOct<ANY_HIT, ORDERED, 1>(sc, r, s, stack, busy, leafThreshold); break;
    case 2: traverseWarpOct<ANY_HIT, ORDERED, 2>(sc, r, s, stack, busy, leafThreshold); break;
    case 3: traverseWarpOct<ANY_HIT, ORDERED, 3>(sc, r, s, stack, busy, leafThreshold); break;
    case 4: traverseWarpOct<ANY_HIT, ORDERED, 4>(sc, r, s, stack, busy, leafThreshold); break;
    case 5: traverseWarpOct<ANY_HIT, ORDERED, 5>(sc, r, s, stack, busy, leafThreshold); break;
    case 6: traverseWarpOct<ANY_HIT, ORDERED, 6>(sc, r, s, stack, busy, leafThreshold); break;
    case 7: traverseWarpOct<ANY_HIT, ORDERED, 7>(sc, r, s, stack, busy, leafThreshold); break;
    default: traverseWarpOct<ANY_HIT, ORDERED, -1>(sc, r, s, stack, busy, leafThreshold); break;
  }
}

// ---------------------------------------------------------------------------------------------
// 4-wide traversal (wide_bvh.cu): same leaves, same order, same tMax at every leaf as the binary walk
// for every ray whose 1/u components are finite; the others (NaN slab products) use the binary tree.
// ---------------------------------------------------------------------------------------------

// Slab test of one child box of a wide node.  OCT as in childTestsOct; OCT = -1 takes min / max per lane.
template <int OCT>
__device__ __forceinline__ void wideChildTest(const float4& b, float zlo, float zhi, const RayPack& p, float tMax,
                                              bool& pass, float& key) {
  const float2 lo = __fmul2_rn(__fadd2_rn(make_float2(b.x, b.y), p.negOxy), p.invXy);   // t0x t0y
  const float2 hi = __fmul2_rn(__fadd2_rn(make_float2(b.z, b.w), p.negOxy), p.invXy);   // t1x t1y
  const float2 zz = __fmul2_rn(__fadd2_rn(make_float2(zlo, zhi), p.negOzz), p.invZz);   // t0z t1z
  float tN, tF;
  if (OCT >= 0) {
    constexpr bool NX = (OCT & 1) != 0, NY = (OCT & 2) != 0, NZ = (OCT & 4) != 0;
    tN = fmaxf(fmaxf(fmaxf(0.0f, NX ? hi.x : lo.x), NY ? hi.y : lo.y), NZ ? zz.y : zz.x);
    tF = fminf(fminf(fminf(tMax, NX ? lo.x : hi.x), NY ? lo.y : hi.y), NZ ? zz.x : zz.y);
  } else {
    tN = fmaxf(fmaxf(fmaxf(0.0f, fminf(lo.x, hi.x)), fminf(lo.y, hi.y)), fminf(zz.x, zz.y));
    tF = fminf(fminf(fminf(tMax, fmaxf(lo.x, hi.x)), fmaxf(lo.y, hi.y)), fmaxf(zz.x, zz.y));
  }
  key = tN;
  pass = tN <= tF;
}

// One wide-node visit: four child box tests; the first passing child (left-first order) is entered, the
// later passing ones are pushed in reverse order with their entry distance (re-validated against the then
// current tMax when popped, exactly like the binary walk).  Returns false when the walk is finished.
// Decodes child K of a compressed node (wide_bvh.cu, k_wide_compress): coordinate = fma(2^23 + q, step, origin'), exact.
// PRMT builds the float 2^23 + q straight from the byte: 0x4B0000qq.
template <int K>
__device__ __forceinline__ void decodeChild(const float4& v1, const float4& v2, float2 stepXy, float2 stepZz, float2 orgXy,
                                            float2 orgZz, float4& b, float& zlo, float& zhi) {
  constexpr unsigned sel = 0x7540u | (unsigned)K;
  const float lox = __uint_as_float(__byte_perm(__float_as_uint(v1.x), 0x4B000000u, sel));
  const float loy = __uint_as_float(__byte_perm(__float_as_uint(v1.y), 0x4B000000u, sel));
  const float loz = __uint_as_float(__byte_perm(__float_as_uint(v1.z), 0x4B000000u, sel));
  const float hix = __uint_as_float(__byte_perm(__float_as_uint(v1.w), 0x4B000000u, sel));
  const float hiy = __uint_as_float(__byte_perm(__float_as_uint(v2.x), 0x4B000000u, sel));
  const float hiz = __uint_as_float(__byte_perm(__float_as_uint(v2.y), 0x4B000000u, sel));
  const float2 lo = __ffma2_rn(make_float2(lox, loy), stepXy, orgXy);
  const float2 hi = __ffma2_rn(make_float2(hix, hiy), stepXy, orgXy);
  const float2 zz = __ffma2_rn(make_float2(loz, hiz), stepZz, orgZz);
  b = make_float4(lo.x, lo.y, hi.x, hi.y);
  zlo = zz.x; zhi = zz.y;
}

template <int OCT, class STK, bool CMP>
__device__ __forceinline__ bool wideStep(const DeviceScene& sc, const RayPack& rp, Trav& s, const STK& stack) {
  YB_CNT(s, kCntWide);
  float4 b0, b1, b2, b3, z01, z23, rf;
  if (CMP) {
    const float4* np = sc.cwide + 4 * (size_t)s.cur;
    const float4 v0 = __ldg(np + 0), v1 = __ldg(np + 1), v2 = __ldg(np + 2);
    rf = __ldg(np + 3);
    const uint32_t eb = __float_as_uint(v0.w);
    const float sx = __uint_as_float((eb & 0xFFu) << 23), sy = __uint_as_float((eb & 0xFF00u) << 15),
                sz = __uint_as_float((eb & 0xFF0000u) << 7);
    const float2 stepXy = make_float2(sx, sy), stepZz = make_float2(sz, sz);
    const float2 orgXy = make_float2(v0.x, v0.y), orgZz = make_float2(v0.z, v0.z);
    decodeChild<0>(v1, v2, stepXy, stepZz, orgXy, orgZz, b0, z01.x, z01.y);
    decodeChild<1>(v1, v2, stepXy, stepZz, orgXy, orgZz, b1, z01.z, z01.w);
    decodeChild<2>(v1, v2, stepXy, stepZz, orgXy, orgZz, b2, z23.x, z23.y);
    decodeChild<3>(v1, v2, stepXy, stepZz, orgXy, orgZz, b3, z23.z, z23.w);
  } else {
    const float4* np = sc.wide + kWideNodeVec * (size_t)s.cur;
    b0 = __ldg(np + 0); b1 = __ldg(np + 1); b2 = __ldg(np + 2); b3 = __ldg(np + 3);
    z01 = __ldg(np + 4); z23 = __ldg(np + 5); rf = __ldg(np + 6);
  }
  bool p0, p1, p2, p3;
  float k0, k1, k2, k3;
  wideChildTest<OCT>(b0, z01.x, z01.y, rp, s.tMax, p0, k0);
  wideChildTest<OCT>(b1, z01.z, z01.w, rp, s.tMax, p1, k1);
  wideChildTest<OCT>(b2, z23.x, z23.y, rp, s.tMax, p2, k2);
  wideChildTest<OCT>(b3, z23.z, z23.w, rp, s.tMax, p3, k3);
  const uint32_t r0 = __float_as_uint(rf.x), r1 = __float_as_uint(rf.y), r2 = __float_as_uint(rf.z),
                 r3 = __float_as_uint(rf.w);
  if (OCT < 0) {       // per-lane min / max cannot tell the inverted box of an empty slot from a real one
    p0 = p0 && r0 != kDevRefNull; p1 = p1 && r1 != kDevRefNull; p2 = p2 && r2 != kDevRefNull; p3 = p3 && r3 != kDevRefNull;
  }
  // branch-free push: a later child goes on the stack iff it passes and an earlier one does too; the stores
  // are unconditional (one slot past the top is scratch), only the stack pointer moves conditionally
  int sp = s.sp;
  const bool e3 = p3 && (p0 || p1 || p2), e2 = p2 && (p0 || p1), e1 = p1 && p0;
  if (e3) stack.store(sp, make_uint2(r3, __float_as_uint(k3)));
  sp += e3 ? 1 : 0;
  if (e2) stack.store(sp, make_uint2(r2, __float_as_uint(k2)));
  sp += e2 ? 1 : 0;
  if (e1) stack.store(sp, make_uint2(r1, __float_as_uint(k1)));
  sp += e1 ? 1 : 0;
#ifdef YB_COUNT_WORK
  s.cnt[kCntPush] += (uint32_t)(sp - s.sp);
#endif
  s.sp = sp;
  if (p0 || p1 || p2 || p3) {
    s.cur = p0 ? r0 : (p1 ? r1 : (p2 ? r2 : r3));
    return true;
  }
  return popNext(s, stack);
}

// Inner phase of the wide walk: wide-node steps until no lane has inner work or enough lanes hold a leaf.
// Only this loop is specialised per octant; the leaf code exists once (traverseWarpWide).
template <int OCT, class STK, bool CMP>
__device__ __forceinline__ void widePhase(const DeviceScene& sc, const RayPack& rp, Trav& s, const STK& stack, bool& run,
                                          int leafThreshold) {
  for (;;) {
    const bool atLeaf = run && (s.cur & kDevRefLeafBit);
    const bool atInner = run && !atLeaf;
    if (!__any_sync(kFull, atInner)) break;
    if (__popc(__ballot_sync(kFull, atLeaf)) >= leafThreshold) break;
    if (atInner) run = wideStep<OCT, STK, CMP>(sc, rp, s, stack);
  }
}

// Compressed walk: the leaf's EXACT box against the ray's current tMax -- the reference's wrapCollider test on entry to
// the leaf (AABBs.hs:42-43) -- before its primitives are tested; the inner boxes on the way down were only supersets.
// (Only rays with finite 1/u walk the wide tree, so min / max are plain.)
__device__ __forceinline__ bool leafBoxPasses(const DeviceScene& sc, const Ray& r, const Trav& s) {
  const float4* bp = ((s.cur & kDevRefMultiBits) == kDevRefMultiBits ? sc.multiBox : sc.leafBox) + 2 * (size_t)(s.cur & 0x3FFFFFFFu);
  const float4 a = __ldg(bp + 0), b = __ldg(bp + 1);
  float key;
  return boxTest(a.x, a.y, a.z, a.w, b.x, b.y, r, s.tMax, key);
}

template <bool ANY_HIT, bool CMP = false, class STK>
__device__ __forceinline__ void traverseWarpWide(const DeviceScene& sc, const Ray& r, Trav& s, const STK& stack, bool busy,
                                                 int leafThreshold, bool leafRun, bool anyRt = false,
                                                 bool octants = true) {
  const bool nanLane = busy && r.exactNaN;
  bool run = busy && !r.exactNaN;
  const unsigned mRun = __ballot_sync(kFull, run);
  if (mRun) {
    int oct = -1;
    const int mine = rayOctant(r);
    const int lead = __shfl_sync(kFull, mine, __ffs(mRun) - 1);
    // octants = false: every warp runs the ONE generic inner loop (min / max per lane): ~6 % more instructions per
    // step, but the kernels that mix closest-hit and any-hit walks of different octants on one SM (k_wf_persist) keep a
    // much smaller set of hot instructions (the L1.5 instruction cache holds 32 KB)
    if (octants && __all_sync(kFull, !run || mine == lead)) oct = lead;
    const RayPack rp = packRay(r);
    for (;;) {
      switch (oct) {
        case 0: widePhase<0, STK, CMP>(sc, rp, s, stack, run, leafThreshold); break;
        case 1: widePhase<1, STK, CMP>(sc, rp, s, stack, run, leafThreshold); break;
        case 2: widePhase<2, STK, CMP>(sc, rp, s, stack, run, leafThreshold); break;
        case 3: widePhase<3, STK, CMP>(sc, rp, s, stack, run, leafThreshold); break;
        case 4: widePhase<4, STK, CMP>(sc, rp, s, stack, run, leafThreshold); break;
        case 5: widePhase<5, STK, CMP>(sc, rp, s, stack, run, leafThreshold); break;
        case 6: widePhase<6, STK, CMP>(sc, rp, s, stack, run, leafThreshold); break;
        case 7: widePhase<7, STK, CMP>(sc, rp, s, stack, run, leafThreshold); break;
        default: widePhase<-1, STK, CMP>(sc, rp, s, stack, run, leafThreshold); break;
      }
      // every running lane now holds a leaf, or enough of them do
      const bool atLeaf = run && (s.cur & kDevRefLeafBit);
      if (!__any_sync(kFull, atLeaf)) break;
      // a lane whose next pending subtree is again a leaf (siblings in one wide node) tests it right away
      if (atLeaf) {
        do {
          if (CMP && !leafBoxPasses(sc, r, s)) run = popNext(s, stack);
          else run = leafStep<ANY_HIT, false>(sc, r, s, stack, anyRt);
        } while (leafRun && run && (s.cur & kDevRefLeafBit));
      }
    }
  }
  if (__any_sync(kFull, nanLane)) traverseWarpOct<ANY_HIT, false, -1>(sc, r, s, stack, nanLane, 1, anyRt);
}

// Counting build: adds this batch's per-lane counters to workStats[kind * 8 + k] (kind 0 = closest-hit walks, 1 = any-hit
// walks); slot 6 of a kind counts the walks themselves, slot 7 the shaded hits.  No-op in the production build.
__device__ __forceinline__ void flushCounts(const WavefrontParams& W, Trav& s, int kind, bool shaded) {
#ifdef YB_COUNT_WORK
  if (!W.workStats) return;
  const unsigned lane = threadIdx.x & 31u;
  for (int k = 0; k < kNumCnt; ++k) {
    const uint32_t v = __reduce_add_sync(kFull, s.cnt[k]);
    if (lane == 0 && v) atomicAdd(&W.workStats[kind * 8 + k], (unsigned long long)v);
    s.cnt[k] = 0u;
  }
  const uint32_t sh = __reduce_add_sync(kFull, shaded ? 1u : 0u);
  if (lane == 0 && sh) atomicAdd(&W.workStats[kind * 8 + 7], (unsigned long long)sh);
#endif
}

// item index (tile-major) -> tile.  Tiles are almost uniform in size, so a proportional guess is
// off by a step or two at most.
__device__ __forceinline__ uint32_t tileOfItem(const WavefrontParams& W, uint32_t item) {
  const uint32_t g = item + W.itemBase;
  uint32_t t = (uint32_t)(((unsigned long long)item * W.base.nTiles) / W.nItems);
  while (__ldg(&W.tileStart[t]) > g) --t;
  while (__ldg(&W.tileStart[t + 1]) <= g) ++t;
  return t;
}

// Pixel of an item.  Inside a tile the items run over strips of 4 rows, column by column, so that
// 32 consecutive items (one warp) cover an 8 x 4 pixel block.  (The reference's own order inside a
// tile, u-major (Sampling.hs:6), only matters for its list layout, not for the image.)
__device__ __forceinline__ void itemPixel(const WavefrontParams& W, uint32_t item, int& u, int& v) {
  if (W.itemPixels) {                      // precomputed once per tile set (k_wf_pixel_table): one coalesced load
    const uint32_t p = __ldg(&W.itemPixels[item + W.itemBase]);
    u = (int)(p & 0xFFFFu); v = (int)(p >> 16);
    return;
  }
  const uint32_t t = tileOfItem(W, item);
  const int4 win = __ldg(&W.base.tiles[t]);
  const int i = (int)(item + W.itemBase - __ldg(&W.tileStart[t]));
  const int tw = win.z - win.x, th = win.w - win.y;
  const int strip = i / (4 * tw);
  const int rem = i - strip * 4 * tw;
  const int rows = min(4, th - 4 * strip);
  u = win.x + rem / rows;
  v = win.y + 4 * strip + rem % rows;
}

template <bool CALL = false>
__device__ __forceinline__ Ray itemRay(const WavefrontParams& W, int u, int v, uint32_t sample) {
  const uint32_t pixel = (uint32_t)(W.base.width * v + u);
  const float fu = (float)u + sampleOffset(W.base.seed, pixel, sample, 0);
  const float fv = (float)v + sampleOffset(W.base.seed, pixel, sample, 1);
  return cameraRay<CALL>(W.base, fu, fv);
}

// Streamed host output: the lanes with `done` have just stored the final value of a pixel of tile row `row`.  One atomic
// per distinct row of the warp (a batch touches one or two); the warp that completes a row publishes it to the host.
// Release pattern: the lanes' stores are ordered before the leader's RELEASE atomic by the warp barrier (cumulativity),
// and the publishing lane fences at system scope.  The release atomic is inline PTX on purpose: __threadfence() is
// fence.sc (MEMBAR.SC.GPU + CCTL.IVALL) and invalidates the SM's whole L1 -- once per batch that cost the walk its
// tree hit rate (+0.3 ms per C4 frame, profiles/r1ab); atom.release.gpu is MEMBAR.ALL.GPU + ATOMG, no invalidation.
__device__ __forceinline__ void rowsSignal(const WavefrontParams& W, uint32_t row, bool done, unsigned lane) {
  __syncwarp();
  unsigned todo = __ballot_sync(kFull, done);
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const uint32_t r = __shfl_sync(kFull, row, leader);
    const unsigned m = __ballot_sync(kFull, done && row == r);
    if ((int)lane == leader) {
      const uint32_t n = (uint32_t)__popc(m);
      uint32_t before;
      asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(before) : "l"(W.rowDone + r), "r"(n) : "memory");
      if (before + n == W.rowItems[r]) {
        __threadfence_system();                    // acquire the other warps' counts, release to the host (128 x per frame)
        W.rowFlags[r] = W.rowSeq;
      }
    }
    todo &= ~m;
  }
}

// JuicyPixels' ImageRGBF -> 8-bit conversion of savePngImage (main.hs:142), as k_quantize_rgb8 (kernels.cu).
__device__ __forceinline__ unsigned char quantize8(float x) {
  const float m = (1.0f <= x) ? 1.0f : x;
  const float c = (0.0f <= m) ? m : 0.0f;
  return (unsigned char)(int)(255.0f * c);
}

// Final value of a pixel from one of the per-batch kernels (every pixel is stored exactly once): the float frame, or
// -- 8-bit host-buffer entry -- the quantised bytes only.
__device__ __forceinline__ void storeFinal(const WavefrontParams& W, size_t index, float x, float y, float z) {
  if (W.rgb8) {
    unsigned char* o = W.rgb8 + 3 * index;
    o[0] = quantize8(x); o[1] = quantize8(y); o[2] = quantize8(z);
  } else {
    float* o = W.sampleOut + 3 * index;
    o[0] = x; o[1] = y; o[2] = z;
  }
}

// The shadow probe of a lane, kept in registers by the fused kernel (k_wf_fused) instead of the queue.
struct LocalProbe {
  V3 origin, dir, contrib;
  float tMax;
  uint32_t nanBits, row;
  bool emit;
};

template <bool AREA, bool FUSED>
__device__ __forceinline__ void shadeAndEmit(const WavefrontParams& W, bool valid, uint32_t item, uint32_t pixel, int v,
                                             const Ray& r, float tHit, uint32_t idx, unsigned lane, uint32_t sLocal,
                                             LocalProbe* local = nullptr) {
  const DeviceScene& sc = W.base.sc;
  const bool hit = valid && idx != kNoHit;
  Surface surf;
  Frame fr;
  MaterialD mat;
  uint32_t nanBits = 0;
  // streamed host output: the pixel's tile row travels with the probe (bits 3.. of q2.w) so that the shadow kernel
  // can count the pixel as finished without a division
  const uint32_t row = (W.rowFlags && valid) ? (uint32_t)W.rowOfV[v] : 0u;
  V3 wo = vneg(r.d);
  if (valid) {
    // the frame of sample sLocal of this launch (sampleOut is the image itself when spp = 1)
    const uint32_t outIndex = sLocal * W.framePixels + pixel;
    float* out = W.sampleOut + 3 * (size_t)outIndex;
    const bool firstSample = W.sample + sLocal == 0u;
    if (!hit) {
      // Nothing -> Vec3 0 0 0.  Final for this pixel unless several slots add to it later (then the float frame it is);
      // storeFinal writes the 8-bit frame instead when the 8-bit host entry streams its rows
      if (FUSED || !(W.dense || W.twoSlot)) storeFinal(W, outIndex, 0.0f, 0.0f, 0.0f);
      else { out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f; }
      if (W.base.primid && firstSample) W.base.primid[pixel] = kNoHit;
    } else {
      surf = surfaceAt<FUSED>(sc, idx, r, tHit);
      if (W.base.primid && firstSample) W.base.primid[pixel] = surf.primId;
      mat = loadMaterial(sc, surf.material);
      fr = makeFrame<FUSED>(surf);
      // ((n . r) @* f r) * rs with rs = vcast 0 = 0 (Integrators.hs:26,37,41-43): +-0, or NaN when the
      // weight is not finite.  Stored as the pixel's base value; the direct term is added to it.
      const V3 refl = vsub(r.d, vscale(2.0f * dot(r.d, surf.n), surf.n));
      const V3 w = vscale(dot(surf.n, refl), bsdfAt<FUSED>(mat, fr, refl, wo));
      const V3 base = vadd(vmul(w, mk(0.0f, 0.0f, 0.0f)), mk(0.0f, 0.0f, 0.0f));     // +0, or NaN
      nanBits = (base.x != base.x ? 1u : 0u) | (base.y != base.y ? 2u : 0u) | (base.z != base.z ? 4u : 0u);
      // several slots: the base is stored now and k_wf_resolve adds to it.  One slot: every pixel is stored exactly
      // ONCE -- here (below) when no probe is emitted, else by the shadow kernel -- so the frame may live in a
      // peer GPU or in mapped host memory without any traffic beyond the frame itself.
      if (W.dense || W.twoSlot) { out[0] = base.x; out[1] = base.y; out[2] = base.z; }
    }
  }
  // one probe per light SLOT with lensq k > 0 (point lights, then every sample of every area light -- the
  // extension of include/yahr_b200.h).  A single slot: compacted queue (warp-aggregated append); several:
  // dense entries item * nSlots + slot so that k_wf_resolve can sum them in slot order.
  uint32_t nEmit = 0;
  auto doSlot = [&](uint32_t slot, V3 lightPos, V3 spectrum, bool area, V3 lightNormal) {
    bool emit = false;
    V3 p0 = mk(0, 0, 0), dl = mk(0, 0, 0), contrib = mk(0, 0, 0);
    if (hit) {
      const V3 pointToLight = vsub(lightPos, surf.x);
      const V3 lightDir = vnorm<FUSED>(pointToLight);
      const V3 k = bsdfAt<FUSED>(mat, fr, lightDir, wo);
      const float cosL = area ? -dot(lightDir, lightNormal) : 1.0f;
      if (lensq(k) > 0.0f && (!area || cosL > 0.0f)) {
        emit = true;
        p0 = vadd(surf.x, vscale(0.001f, lightDir));
        dl = vsub(lightPos, p0);
        const V3 intensity = vscale(rcp(lensq(pointToLight)), area ? vscale(cosL, spectrum) : spectrum);
        contrib = vmul(vscale(fabsf(dot(lightDir, surf.n)), k), intensity);
      }
    }
    if (FUSED) {                                   // one slot, probe stays in registers
      if (emit) {
        local->origin = p0; local->dir = vnorm<FUSED>(dl); local->tMax = len(dl); local->contrib = contrib;
        local->emit = true;
        ++nEmit;
      }
      return;
    }
    uint32_t e = 0;
    if (W.dense) {
      e = (sLocal * W.nItems + item) * sc.nSlots + slot;
      if (valid && !emit) W.q0[e] = make_float4(0.0f, 0.0f, 0.0f, -1.0f);     // empty slot
    } else {
      const unsigned m = __ballot_sync(kFull, emit);
      if (m) {
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(&W.work[2], (uint32_t)__popc(m));
        base = __shfl_sync(kFull, base, leader);
        e = base + __popc(m & ((1u << lane) - 1u));
      }
    }
    if (emit) {
      const V3 d = vnorm(dl);
      W.q0[e] = make_float4(p0.x, p0.y, p0.z, len(dl));          // probe origin, tMax = len (p1 - p0)
      W.q1[e] = make_float4(d.x, d.y, d.z, __uint_as_float(sLocal * W.framePixels + pixel));
      W.q2[e] = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(nanBits | (row << 3)));
      ++nEmit;
    }
  };
  uint32_t slot = 0;
  for (uint32_t li = 0; li < sc.nLights; ++li, ++slot)
    doSlot(slot, xyz(__ldg(&sc.lights[2 * li + 0])), xyz(__ldg(&sc.lights[2 * li + 1])), false, mk(0, 0, 0));
  if (AREA) {
    ShadeCtx ctx;
    ctx.seed = W.base.seed; ctx.pixel = pixel; ctx.sample = W.sample + sLocal; ctx.level = 0;
    for (uint32_t a = 0; a < sc.nAreaLights; ++a) {
      const AreaLightD l = loadAreaLight(sc, a);
      for (uint32_t j = 0; j < l.samples; ++j, ++slot) doSlot(slot, areaLightPoint(l, ctx, slot), l.flux, true, l.normal);
    }
  }
  if (hit && !W.dense && !W.twoSlot && nEmit == 0u) {
    const float qnan = __uint_as_float(0x7FFFFFFFu);
    const size_t index = (size_t)(sLocal * W.framePixels + pixel);
    const float x = (nanBits & 1u) ? qnan : 0.0f, y = (nanBits & 2u) ? qnan : 0.0f, z = (nanBits & 4u) ? qnan : 0.0f;
    storeFinal(W, index, x, y, z);
  }
  if (FUSED) { local->nanBits = nanBits; local->row = row; }
  else if (W.rowFlags) rowsSignal(W, row, valid && nEmit == 0u, lane);     // pixels that are final without a shadow probe
  const uint32_t warpEmit = __reduce_add_sync(kFull, nEmit);    // shadow-ray count for the stats
  if (lane == 0 && warpEmit) atomicAdd(&W.work[3], warpEmit);
}

// Result of one shadow probe.  Single slot: the shading step left the pixel to this kernel, which stores it exactly
// once: the base value (+0, or NaN: flag bits in q2.w) for an occluded probe, base + (0 + contribution) for an
// unoccluded one -- no read-modify-write and no second store, so the frame may live in a peer GPU or in mapped
// host memory.  Several slots: only the visibility flag is recorded.
__device__ __forceinline__ uint32_t shadowResult(const WavefrontParams& W, uint32_t entry, bool unoccluded) {
  if (W.visibility) { W.visibility[entry] = unoccluded ? 1 : 0; return 0u; }
  if (W.twoSlot) {
    // two light slots: sum = foldl (+) 0 over the visible ones (Integrators.hs:50-61) on top of the base value the
    // shading step stored; (0 + a) + b and (0 + b) + a are the same float, so the two probes of a pixel may land in
    // either order
    if (unoccluded) {
      const float4 c = W.q2[entry];
      float* out = W.sampleOut + 3 * (size_t)__float_as_uint(W.q1[entry].w);
      atomicAdd(out + 0, c.x); atomicAdd(out + 1, c.y); atomicAdd(out + 2, c.z);
    }
    return 0u;
  }
  const float4 b = W.q1[entry];
  float4 c = make_float4(0.0f, 0.0f, 0.0f, W.q2[entry].w);
  if (unoccluded) c = W.q2[entry];
  const uint32_t nanBits = __float_as_uint(c.w) & 7u;
  const float qnan = __uint_as_float(0x7FFFFFFFu);
  storeFinal(W, (size_t)__float_as_uint(b.w), (nanBits & 1u) ? qnan : 0.0f + c.x, (nanBits & 2u) ? qnan : 0.0f + c.y,
             (nanBits & 4u) ? qnan : 0.0f + c.z);
  return __float_as_uint(c.w) >> 3;              // the pixel's tile row (streamed host output)
}

}  // namespace

// ---------------------------------------------------------------------------------------------
template <bool ORDERED, int MIN_BLOCKS, bool WIDE, bool AREA, int SH, bool CMP = false>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_wf_primary(const __grid_constant__ WavefrontParams W) {
  YB_STACK(SH);
  const unsigned lane = threadIdx.x & 31u;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&W.work[0], 32u);
    base = __shfl_sync(kFull, base, 0);
    if (base >= W.itemsPadded * W.samplesPerLaunch) break;
    // work is sample-major: a batch never straddles two samples (itemsPadded is a multiple of 32)
    const uint32_t sLocal = W.samplesPerLaunch > 1u ? base / W.itemsPadded : 0u;
    const uint32_t item = base - sLocal * W.itemsPadded + lane;
    const bool valid = item < W.nItems;
    int u = 0, v = 0;
    Ray r;
    Trav s;
    s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
    YB_CNT_INIT(s);
    bool busy = false;
    if (valid) {
      itemPixel(W, item, u, v);
      r = itemRay(W, u, v, W.sample + sLocal);
      busy = travBegin(W.base.sc, r, 1e6f, s);
    } else {
      r = makeRay(mk(0, 0, 0), mk(0, 0, 1));
    }
    if (WIDE) traverseWarpWide<false, CMP>(W.base.sc, r, s, stack, busy, (int)W.leafThreshold, W.leafRun != 0);
    else traverseWarp<false, ORDERED>(W.base.sc, r, s, stack, busy, (int)W.leafThreshold, W.packed != 0);
    flushCounts(W, s, 0, valid && s.best != kNoHit);
    shadeAndEmit<AREA, false>(W, valid, item, (uint32_t)(W.base.width * v + u), v, r, s.tMax, s.best, lane, sLocal);
  }
}

// ---------------------------------------------------------------------------------------------
// Three-kernel set (W.split): the closest-hit walk and the shading of k_wf_primary as two kernels.
//   k_wf_trace : persistent warps, camera ray + closest-hit walk only; writes (DFS position, t) per item.  No shading
//                code in the kernel: the hot instructions are the walk loops and the leaf test (the L1.5 instruction
//                cache holds 32 KB; k_wf_primary is 67 KB of code and 1.2 of its 8 warps per scheduler wait for
//                instructions, profiles/r2d), and 48 registers instead of 64 let 10 CTAs per SM be resident.
//   k_wf_shade : one thread per item, full warps of straight-line code: regenerates the camera ray with the same
//                instructions (same bits), then shadeAndEmit exactly as k_wf_primary does after its walk.
// Same arithmetic, same queue, same k_wf_shadow afterwards: bit-identical frames.
template <int MIN_BLOCKS, bool CMP>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_wf_trace(const __grid_constant__ WavefrontParams W) {
  YB_STACK(0);
  const unsigned lane = threadIdx.x & 31u;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&W.work[0], 32u);
    base = __shfl_sync(kFull, base, 0);
    if (base >= W.itemsPadded * W.samplesPerLaunch) break;
    const uint32_t sLocal = W.samplesPerLaunch > 1u ? base / W.itemsPadded : 0u;
    const uint32_t item = base - sLocal * W.itemsPadded + lane;
    const bool valid = item < W.nItems;
    Ray r = makeRay(mk(0, 0, 0), mk(0, 0, 1));
    Trav s;
    s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
    YB_CNT_INIT(s);
    bool busy = false;
    if (valid) {
      int u, v;
      itemPixel(W, item, u, v);
      r = itemRay(W, u, v, W.sample + sLocal);
      busy = travBegin(W.base.sc, r, 1e6f, s);
    }
    traverseWarpWide<false, CMP>(W.base.sc, r, s, stack, busy, (int)W.leafThreshold, W.leafRun != 0);
    flushCounts(W, s, 0, valid && s.best != kNoHit);
    W.hits[base + lane] = make_uint2(s.best, __float_as_uint(s.tMax));
  }
}

template <bool AREA>
__global__ void __launch_bounds__(128) k_wf_shade(const __grid_constant__ WavefrontParams W) {
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;              // (sample of the launch, padded item)
  if (idx - lane >= W.itemsPadded * W.samplesPerLaunch) return;            // whole warps only: shadeAndEmit votes
  const uint32_t sLocal = W.samplesPerLaunch > 1u ? idx / W.itemsPadded : 0u;
  const uint32_t item = idx - sLocal * W.itemsPadded;
  const bool valid = item < W.nItems;
  int u = 0, v = 0;
  Ray r = makeRay(mk(0, 0, 0), mk(0, 0, 1));
  uint2 h = make_uint2(kNoHit, 0u);
  if (valid) {
    itemPixel(W, item, u, v);
    r = itemRay(W, u, v, W.sample + sLocal);
    h = W.hits[idx];
  }
  shadeAndEmit<AREA, false>(W, valid, item, (uint32_t)(W.base.width * v + u), v, r, __uint_as_float(h.y), h.x, lane, sLocal);
}

#ifndef YB_COUNT_WORK
// ---------------------------------------------------------------------------------------------
// Fused variant for ONE light slot (the reference's configuration: one point light): the warp that traced and shaded
// a batch walks the shadow probes of that batch itself, straight from registers -- no probe queue, no second kernel,
// every pixel stored once, final, in item order.  Same arithmetic as k_wf_primary + k_wf_shadow (bit-identical
// frames); what changes is the schedule: the probes of a batch are not compacted with those of other batches (lanes
// without a probe idle during the any-hit walk: slower than the two-kernel set on a device-resident frame,
// profiles/r1ab), but pixels become final batch by batch, so the host-buffer entry can stream finished tile rows to
// the host while the rest of the frame is traced (rowsSignal).  Compacting the probes per warp first (a 64-entry ring
// per warp, in shared memory or in global memory around L1) was measured slower still (profiles/r1ab).
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_wf_fused(const __grid_constant__ WavefrontParams W) {
  YB_STACK(0);
  const unsigned lane = threadIdx.x & 31u;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&W.work[0], 32u);
    base = __shfl_sync(kFull, base, 0);
    if (base >= W.itemsPadded * W.samplesPerLaunch) break;
    const uint32_t sLocal = W.samplesPerLaunch > 1u ? base / W.itemsPadded : 0u;
    const uint32_t item = base - sLocal * W.itemsPadded + lane;
    const bool valid = item < W.nItems;
    int u = 0, v = 0;
    Ray r;
    Trav s;
    s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
    bool busy = false;
    if (valid) {
      itemPixel(W, item, u, v);
      r = itemRay<true>(W, u, v, W.sample + sLocal);
      busy = travBegin(W.base.sc, r, 1e6f, s);
    } else {
      r = makeRay(mk(0, 0, 0), mk(0, 0, 1));
    }
    // Both walks of a batch -- closest hit, then the any-hit walk of its probes (reachable, Rays.hs:49-54) -- go through
    // ONE copy of the traversal code (any-hit is a run-time flag of the leaf step; the loop below is not unrolled): with
    // two inlined copies the kernel was 6400 SASS instructions and instruction fetch was its top stall (ncu:
    // stalled_no_instruction 5.1 per issue against 1.1 in k_wf_primary, profiles/r1ac).
    LocalProbe pr;
    pr.emit = false; pr.nanBits = 0; pr.row = 0; pr.tMax = 0.0f;
    pr.origin = mk(0, 0, 0); pr.dir = mk(0, 0, 1); pr.contrib = mk(0, 0, 0);
    const uint32_t pixel = (uint32_t)(W.base.width * v + u);
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      traverseWarpWide<false>(W.base.sc, r, s, stack, busy, (int)W.leafThreshold, W.leafRun != 0, phase == 1);
      if (phase == 0) {
        shadeAndEmit<false, true>(W, valid, item, pixel, v, r, s.tMax, s.best, lane, sLocal, &pr);
        if (!__any_sync(kFull, pr.emit)) break;
        r = makeRay(pr.origin, pr.dir);
        s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
        busy = pr.emit && travBegin(W.base.sc, r, pr.tMax, s);
      } else if (pr.emit) {
        const bool unoccluded = s.best == kNoHit;
        const float qnan = __uint_as_float(0x7FFFFFFFu);
        storeFinal(W, (size_t)(sLocal * W.framePixels + pixel),
                   (pr.nanBits & 1u) ? qnan : 0.0f + (unoccluded ? pr.contrib.x : 0.0f),
                   (pr.nanBits & 2u) ? qnan : 0.0f + (unoccluded ? pr.contrib.y : 0.0f),
                   (pr.nanBits & 4u) ? qnan : 0.0f + (unoccluded ? pr.contrib.z : 0.0f));
      }
    }
    if (W.rowFlags) rowsSignal(W, pr.row, valid, lane);
  }
}

#endif   // !YB_COUNT_WORK

// ---------------------------------------------------------------------------------------------
// ONE persistent kernel per frame for one light slot (the reference's configuration), with COMMITTED PROBE CHUNKS.
// Every warp loops: (1) if its CTA holds a complete chunk of 32 shadow probes, claim it and run the any-hit walks
// (reachable, Rays.hs:49-54) of those 32 probes -- compacted across the CTA's batches, at k_wf_shadow's lane
// occupancy; (2) else take a batch of 32 primary items from the global counter: camera rays, closest-hit walk, shading
// (as k_wf_primary), append the probes of the batch to the CTA's ring (one shared-memory atomic per warp) and COMMIT
// them: add the number of entries written to the counter of each chunk the append touched.  A chunk is claimed only
// once its counter says it is complete, so nobody waits for a producer; the partial chunk a CTA is left with at the
// end is walked once every warp of the CTA has run out of primary work.
//   * one launch: one ramp-up and one tail per frame, and the tail of the primary trace (the long horizon rays)
//     overlaps the any-hit walks -- what a 1/N share of a frame on N GPUs needs most;
//   * the queue is a 256-entry ring PER CTA (32 B per probe: origin + pixel, contribution + flags; direction and length
//     are recomputed from the light): control words in shared memory (claim / commit cost a shared atomic, not an L2
//     round trip -- a first version with ONE global queue serialised on its head counter: 3 us per claim, 411 ms per
//     frame, profiles/r2a), data in global memory that never leaves L2 (9.7 MB for the whole grid, rewritten in place);
//   * a probe is walked by the SM that shaded it, soon after: its walk starts in tree nodes that are still in L1;
//   * pixels become final soon after their batch, roughly in item order: the streamed host rows work as with k_wf_fused;
//   * both walks go through ONE copy of the traversal code (any-hit is a run-time flag), as in k_wf_fused.
// Ring safety: a warp takes primary work only when its CTA has no complete unclaimed chunk, so at most 31 + 4 x 32
// entries are ever pending or being walked; the ring holds 256.  The append still checks that the slot it is about to
// reuse has been released (never observed to wait).
constexpr uint32_t kRingEntries = 256u, kRingChunks = kRingEntries / 32u;

struct PersistCtl {
  uint32_t tail;                 // entries reserved so far (monotonic)
  uint32_t head;                 // chunks claimed so far (monotonic)
  uint32_t producing;            // warps of the CTA between taking a primary batch and committing its probes
  uint32_t noPrimary;            // warps of the CTA that have seen the global primary counter run out
  uint32_t commit[kRingChunks];  // entries written into the chunk currently occupying the ring slot
  uint32_t inUse[kRingChunks];   // the ring slot is being walked
};

template <int MIN_BLOCKS, int SH>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_wf_persist(const __grid_constant__ WavefrontParams W) {
  YB_STACK(SH);
  __shared__ PersistCtl ctl;
  __shared__ float4 stash[128];          // per lane: contribution + flags of the probe being walked
  const DeviceScene& sc = W.base.sc;
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t nWork = W.itemsPadded * W.samplesPerLaunch;
  constexpr uint32_t kNone = 0xFFFFFFFFu;
  volatile PersistCtl* vc = &ctl;
  if (threadIdx.x < sizeof(PersistCtl) / 4u) reinterpret_cast<uint32_t*>(&ctl)[threadIdx.x] = 0u;
  __syncthreads();
  float4* const ring0 = W.q0 + (size_t)blockIdx.x * kRingEntries;
  float4* const ring2 = W.q2 + (size_t)blockIdx.x * kRingEntries;
  bool primaryLeft = true;
  uint32_t idle = 0;
  for (;;) {
    // ---- what this warp does next --------------------------------------------------------------------------------
    uint32_t chunk = kNone, count = 0, base = kNone, fin = 0;
    if (lane == 0) {
      for (;;) {                                                   // (1) a complete chunk of this CTA?
        const uint32_t h = vc->head;
        uint32_t n = 0;
        if (vc->commit[h & (kRingChunks - 1u)] == 32u) n = 32u;
        else if (vc->noPrimary == 4u && vc->producing == 0u) {     // the CTA's queue is final: its last, partial chunk
          const uint32_t t = vc->tail;
          if (32u * h < t && vc->commit[h & (kRingChunks - 1u)] == t - 32u * h) n = t - 32u * h;
        }
        if (n == 0) break;
        if (atomicCAS(&ctl.head, h, h + 1u) == h) {
          chunk = h; count = n;
          vc->inUse[h & (kRingChunks - 1u)] = 1u;
          vc->commit[h & (kRingChunks - 1u)] = 0u;
          break;
        }
      }
      if (chunk == kNone && primaryLeft) {                         // (2) a primary batch
        atomicAdd(&ctl.producing, 1u);
        base = atomicAdd(&W.work[0], 32u);
        if (base >= nWork) { atomicSub(&ctl.producing, 1u); atomicAdd(&ctl.noPrimary, 1u); }
      }
      if (chunk == kNone && (base == kNone || base >= nWork))      // (3) nothing right now: is the CTA finished?
        fin = (vc->noPrimary == 4u && vc->producing == 0u && 32u * vc->head >= vc->tail) ? 1u : 0u;
      __threadfence_block();
    }
    chunk = __shfl_sync(kFull, chunk, 0);
    count = __shfl_sync(kFull, count, 0);
    base = __shfl_sync(kFull, base, 0);
    const bool anyPhase = chunk != kNone;
    if (!anyPhase && (base == kNone || base >= nWork)) {
      primaryLeft = false;
      if (__shfl_sync(kFull, fin, 0)) break;
      // (safety net, never reached in a correct run: ~3 s of idling ends the warp and raises work[6] instead of
      // hanging the device)
      if (++idle > (1u << 24)) { if (lane == 0) atomicExch(&W.work[6], 1u); break; }
      __nanosleep(100);
      continue;
    }
    // ---- set up the walk: a probe of the chunk, or the camera ray of an item ----------------------------------------
    Ray r = makeRay(mk(0, 0, 0), mk(0, 0, 1));
    Trav s;
    s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
    YB_CNT_INIT(s);
    bool busy = false, valid = false;
    int u = 0, v = 0;
    uint32_t sLocal = 0, item = 0, outIndex = 0;
    const uint32_t slot = (32u * chunk + lane) & (kRingEntries - 1u);
    if (anyPhase) {
      valid = lane < count;
      float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (valid) {
        a = __ldcg(&ring0[slot]);
        stash[threadIdx.x] = __ldcg(&ring2[slot]);
      }
      // The ring slot is released as soon as its 32 entries have been READ (the ballot consumes the loaded words, the
      // shared store the other record), not when the walks end: walk times are heavy-tailed, and a slot pinned by one
      // long chunk stalled the CTA's three other warps at the ring's wrap-around (first version: C4 1.89 ms, r2b).
      __ballot_sync(kFull, __float_as_uint(a.w) == 0xFFFFFFFFu);
      if (lane == 0) { __threadfence_block(); vc->inUse[chunk & (kRingChunks - 1u)] = 0u; }
      if (valid) {
        outIndex = __float_as_uint(a.w);
        // illuminationAtPoint's probe (Lights.hs:20-24) from its origin: direction and length as shadeAndEmit computes them
        const V3 p0 = mk(a.x, a.y, a.z);
        const V3 dl = vsub(xyz(__ldg(&sc.lights[0])), p0);
        r = makeRay(p0, vnorm<true>(dl));
        busy = travBegin(sc, r, len(dl), s);
      }
    } else {
      sLocal = W.samplesPerLaunch > 1u ? base / W.itemsPadded : 0u;
      item = base - sLocal * W.itemsPadded + lane;
      valid = item < W.nItems;
      if (valid) {
        itemPixel(W, item, u, v);
        r = itemRay<true>(W, u, v, W.sample + sLocal);
        busy = travBegin(sc, r, 1e6f, s);
      }
    }
    traverseWarpWide<false>(sc, r, s, stack, busy, (int)W.leafThreshold, W.leafRun != 0, anyPhase, W.packed != 0);
    flushCounts(W, s, anyPhase ? 1 : 0, !anyPhase && valid && s.best != kNoHit);
    if (anyPhase) {
      // ---- result of the probes: every pixel of the chunk is stored exactly once, here ------------------------------
      uint32_t row = 0;
      if (valid) {
        const float4 c = stash[threadIdx.x];
        const bool unoccluded = s.best == kNoHit;
        const uint32_t flags = __float_as_uint(c.w);
        const float qnan = __uint_as_float(0x7FFFFFFFu);
        storeFinal(W, (size_t)outIndex, (flags & 1u) ? qnan : 0.0f + (unoccluded ? c.x : 0.0f),
                   (flags & 2u) ? qnan : 0.0f + (unoccluded ? c.y : 0.0f),
                   (flags & 4u) ? qnan : 0.0f + (unoccluded ? c.z : 0.0f));
        row = flags >> 3;
      }
      if (W.rowFlags) rowsSignal(W, row, valid, lane);
    } else {
      // ---- shading; the probes of the batch go to the CTA's ring -----------------------------------------------------
      LocalProbe pr;
      pr.emit = false; pr.nanBits = 0; pr.row = 0; pr.tMax = 0.0f;
      pr.origin = mk(0, 0, 0); pr.dir = mk(0, 0, 1); pr.contrib = mk(0, 0, 0);
      const uint32_t pixel = (uint32_t)(W.base.width * v + u);
      shadeAndEmit<false, true>(W, valid, item, pixel, v, r, s.tMax, s.best, lane, sLocal, &pr);
      const unsigned m = __ballot_sync(kFull, pr.emit);
      const uint32_t n = (uint32_t)__popc(m);
      uint32_t first = 0;
      if (lane == 0 && n) {
        first = atomicAdd(&ctl.tail, n);
        // the ring slots about to be written must have been released: chunk k reuses the slot of chunk k - 8
        const uint32_t lastChunk = (first + n - 1u) >> 5;
        uint32_t spins = 0;
        while ((lastChunk >= kRingChunks && vc->head + kRingChunks <= lastChunk) ||
               vc->inUse[lastChunk & (kRingChunks - 1u)] || vc->inUse[(first >> 5) & (kRingChunks - 1u)]) {
          if (vc->head > lastChunk) break;                         // (cannot happen: our chunk is not committed yet)
          if (++spins > (1u << 24)) { atomicExch(&W.work[6], 2u); break; }
          __nanosleep(50);
        }
      }
      first = __shfl_sync(kFull, first, 0);
      if (pr.emit) {
        const uint32_t e = (first + (uint32_t)__popc(m & ((1u << lane) - 1u))) & (kRingEntries - 1u);
        __stcg(&ring0[e], make_float4(pr.origin.x, pr.origin.y, pr.origin.z, __uint_as_float(sLocal * W.framePixels + pixel)));
        __stcg(&ring2[e], make_float4(pr.contrib.x, pr.contrib.y, pr.contrib.z, __uint_as_float(pr.nanBits | (pr.row << 3))));
      }
      __syncwarp();
      if (lane == 0) {                                             // commit: the append touches one or two chunks
        __threadfence_block();
        if (n) {
          const uint32_t c0 = first >> 5, inFirst = 32u - (first & 31u);
          if (n <= inFirst) atomicAdd(&ctl.commit[c0 & (kRingChunks - 1u)], n);
          else {
            atomicAdd(&ctl.commit[c0 & (kRingChunks - 1u)], inFirst);
            atomicAdd(&ctl.commit[(c0 + 1u) & (kRingChunks - 1u)], n - inFirst);
          }
        }
        __threadfence_block();
        atomicSub(&ctl.producing, 1u);
      }
      if (W.rowFlags) rowsSignal(W, pr.row, valid && !pr.emit, lane);            // final without a probe
    }
  }
}

#ifndef YB_COUNT_WORK
// ---------------------------------------------------------------------------------------------
// Whitted recursion (recursionDepth >= 2, Integrators.hs:22-47) for ONE point light, on the same plan as k_wf_fused: a
// warp takes a batch through level after level -- closest-hit walk, shading, any-hit walk of the level's shadow probe,
// then the mirror ray r = u - 2 (u . n) n from x + 0.001 r -- with ONE copy of the walk code for all of them, and
// folds the levels at the end exactly as the megakernel does: acc = weight[k] * acc + direct[k] from the deepest level
// up, weight = (n . r) @* f r, direct = 0 + contribution (sum = foldl (+) 0), a miss ends the path (Nothing -> 0).
// The scene shipped with the reference uses depth 3, its Blender exporter depth 2 (scene.yahrr:2, render_engine.py:65).
struct LevelShade {
  V3 weight, nextOrigin, nextDir;          // ((n . r) @* f r), mirror ray
  V3 probeOrigin, probeDir, contrib;       // shadow probe of the point light and what it contributes when unoccluded
  float probeTMax;
  bool hit, emit;
};

__device__ __forceinline__ void shadeLevel(const WavefrontParams& W, bool alive, const Ray& r, float tHit, uint32_t idx,
                                           bool writeId, uint32_t pixel, LevelShade& o) {
  const DeviceScene& sc = W.base.sc;
  o.hit = alive && idx != kNoHit;
  o.emit = false;
  if (!o.hit) {
    if (alive && writeId && W.base.primid) W.base.primid[pixel] = kNoHit;
    return;
  }
  const Surface surf = surfaceAt<true>(sc, idx, r, tHit);
  if (writeId && W.base.primid) W.base.primid[pixel] = surf.primId;
  const MaterialD mat = loadMaterial(sc, surf.material);
  const Frame fr = makeFrame<true>(surf);
  const V3 wo = vneg(r.d);
  const V3 refl = vsub(r.d, vscale(2.0f * dot(r.d, surf.n), surf.n));          // reflectionDir
  o.weight = vscale(dot(surf.n, refl), bsdfAt<true>(mat, fr, refl, wo));
  o.nextOrigin = vadd(surf.x, vscale(0.001f, refl));
  o.nextDir = refl;
  // directIllumination / illuminationAtPoint for the one point light (same expressions as shadeAndEmit's slot)
  const V3 lightPos = xyz(__ldg(&sc.lights[0])), spectrum = xyz(__ldg(&sc.lights[1]));
  const V3 pointToLight = vsub(lightPos, surf.x);
  const V3 lightDir = vnorm<true>(pointToLight);
  const V3 k = bsdfAt<true>(mat, fr, lightDir, wo);
  if (lensq(k) > 0.0f) {
    o.emit = true;
    o.probeOrigin = vadd(surf.x, vscale(0.001f, lightDir));
    const V3 dl = vsub(lightPos, o.probeOrigin);
    o.probeDir = vnorm<true>(dl);
    o.probeTMax = len(dl);
    const V3 intensity = vscale(rcp(lensq(pointToLight)), spectrum);
    o.contrib = vmul(vscale(fabsf(dot(lightDir, surf.n)), k), intensity);
  }
}

constexpr int kMaxLevels = 16;             // YAHR_B200_MAX_RECURSION

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_wf_fused_depth(const __grid_constant__ WavefrontParams W) {
  YB_STACK(0);
  V3 weight[kMaxLevels], direct[kMaxLevels];
  const unsigned lane = threadIdx.x & 31u;
  const int depth = W.base.depth;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&W.work[0], 32u);
    base = __shfl_sync(kFull, base, 0);
    if (base >= W.itemsPadded * W.samplesPerLaunch) break;
    const uint32_t sLocal = W.samplesPerLaunch > 1u ? base / W.itemsPadded : 0u;
    const uint32_t item = base - sLocal * W.itemsPadded + lane;
    const bool valid = item < W.nItems;
    int u = 0, v = 0;
    Ray r = makeRay(mk(0, 0, 0), mk(0, 0, 1));
    Trav s;
    s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
    bool busy = false;
    if (valid) {
      itemPixel(W, item, u, v);
      r = itemRay<true>(W, u, v, W.sample + sLocal);
      busy = travBegin(W.base.sc, r, 1e6f, s);
    }
    const uint32_t pixel = (uint32_t)(W.base.width * v + u);
    const bool firstSample = W.sample + sLocal == 0u;
    bool alive = valid;                    // the path of this lane still continues
    int levels = 0;                        // levels recorded in weight[] / direct[]
    uint32_t nProbes = 0, nSecondary = 0;
    LevelShade ls;
    ls.hit = false; ls.emit = false; ls.probeTMax = 0.0f;
    ls.weight = ls.nextOrigin = ls.probeOrigin = ls.contrib = mk(0, 0, 0);
    ls.nextDir = ls.probeDir = mk(0, 0, 1);
#pragma unroll 1
    for (int step = 0; step < 2 * depth; ++step) {
      const bool anyPhase = (step & 1) != 0;
      const int level = step >> 1;
      if (!anyPhase && !__any_sync(kFull, alive)) break;
      traverseWarpWide<false>(W.base.sc, r, s, stack, busy, (int)W.leafThreshold, W.leafRun != 0, anyPhase);
      if (!anyPhase) {
        shadeLevel(W, alive, r, s.tMax, s.best, level == 0 && firstSample, pixel, ls);
        alive = ls.hit;
        if (alive) weight[level] = ls.weight;
        if (ls.emit) { r = makeRay(ls.probeOrigin, ls.probeDir); ++nProbes; }
        s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
        busy = ls.emit && travBegin(W.base.sc, r, ls.probeTMax, s);
      } else {
        if (alive) {
          const bool lit = ls.emit && s.best == kNoHit;
          direct[level] = vadd(mk(0.0f, 0.0f, 0.0f), lit ? ls.contrib : mk(0.0f, 0.0f, 0.0f));
          levels = level + 1;
        }
        alive = alive && level + 1 < depth;
        s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
        busy = false;
        if (alive) {
          r = makeRay(ls.nextOrigin, ls.nextDir);
          busy = travBegin(W.base.sc, r, 1e6f, s);
          ++nSecondary;
        }
      }
    }
    if (valid) {
      V3 acc = mk(0.0f, 0.0f, 0.0f);
      for (int k = levels - 1; k >= 0; --k) acc = vadd(vmul(weight[k], acc), direct[k]);
      storeFinal(W, (size_t)(sLocal * W.framePixels + pixel), acc.x, acc.y, acc.z);
    }
    if (W.rowFlags) rowsSignal(W, valid ? (uint32_t)W.rowOfV[v] : 0u, valid, lane);       // streamed host output
    const uint32_t wp = __reduce_add_sync(kFull, nProbes), ws = __reduce_add_sync(kFull, nSecondary);
    if (lane == 0 && wp) atomicAdd(&W.work[3], wp);
    if (lane == 0 && ws) atomicAdd(&W.work[4], ws);
  }
}

// The same for SEVERAL point lights: a level is one closest-hit walk and one any-hit walk per light, in light order
// (directIllumination's sum = foldl (+) 0 over the lights, Integrators.hs:50-61).  Between the walks of a level only
// (ray, primitive, t) of the hit is kept; the surface, frame and material are rebuilt from them before each probe.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_wf_fused_depth_lights(const __grid_constant__ WavefrontParams W) {
  YB_STACK(0);
  V3 weight[kMaxLevels], direct[kMaxLevels];
  const unsigned lane = threadIdx.x & 31u;
  const DeviceScene& sc = W.base.sc;
  const int depth = W.base.depth;
  const int nL = (int)sc.nLights, per = 1 + nL;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&W.work[0], 32u);
    base = __shfl_sync(kFull, base, 0);
    if (base >= W.itemsPadded * W.samplesPerLaunch) break;
    const uint32_t sLocal = W.samplesPerLaunch > 1u ? base / W.itemsPadded : 0u;
    const uint32_t item = base - sLocal * W.itemsPadded + lane;
    const bool valid = item < W.nItems;
    int u = 0, v = 0;
    Ray r = makeRay(mk(0, 0, 0), mk(0, 0, 1));
    Trav s;
    s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
    bool busy = false;
    if (valid) {
      itemPixel(W, item, u, v);
      r = itemRay<true>(W, u, v, W.sample + sLocal);
      busy = travBegin(sc, r, 1e6f, s);
    }
    const uint32_t pixel = (uint32_t)(W.base.width * v + u);
    const bool firstSample = W.sample + sLocal == 0u;
    bool alive = valid, emit = false;
    int levels = 0;
    uint32_t nProbes = 0, nSecondary = 0, hitIdx = kNoHit;
    float hitT = 0.0f;
    V3 lo = r.o, ld = r.d;                       // the ray of the current level
    V3 total = mk(0, 0, 0), contrib = mk(0, 0, 0);
#pragma unroll 1
    for (int step = 0; step < depth * per; ++step) {
      const int level = step / per, phase = step - level * per;      // phase 0: closest hit; 1..nL: probe of light phase-1
      if (phase == 0 && !__any_sync(kFull, alive)) break;
      traverseWarpWide<false>(sc, r, s, stack, busy, (int)W.leafThreshold, W.leafRun != 0, phase != 0);
      if (phase == 0) {
        const bool hit = alive && s.best != kNoHit;
        if (alive && !hit && level == 0 && firstSample && W.base.primid) W.base.primid[pixel] = kNoHit;
        alive = hit; hitIdx = s.best; hitT = s.tMax;
        total = mk(0.0f, 0.0f, 0.0f);
      } else if (alive) {
        const bool lit = emit && s.best == kNoHit;
        total = vadd(total, lit ? contrib : mk(0.0f, 0.0f, 0.0f));
      }
      // what the next walk of this lane is: the probe of light `phase`, or the mirror ray of the next level
      emit = false; busy = false;
      s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
      if (alive) {
        const Ray lr = makeRay(lo, ld);
        const Surface surf = surfaceAt<true>(sc, hitIdx, lr, hitT);
        const MaterialD mat = loadMaterial(sc, surf.material);
        const Frame fr = makeFrame<true>(surf);
        const V3 wo = vneg(ld);
        const V3 refl = vsub(ld, vscale(2.0f * dot(ld, surf.n), surf.n));          // reflectionDir
        if (phase == 0) {
          if (level == 0 && firstSample && W.base.primid) W.base.primid[pixel] = surf.primId;
          weight[level] = vscale(dot(surf.n, refl), bsdfAt<true>(mat, fr, refl, wo));
        }
        if (phase < nL) {
          const V3 lightPos = xyz(__ldg(&sc.lights[2 * phase + 0])), spectrum = xyz(__ldg(&sc.lights[2 * phase + 1]));
          const V3 pointToLight = vsub(lightPos, surf.x);
          const V3 lightDir = vnorm<true>(pointToLight);
          const V3 k = bsdfAt<true>(mat, fr, lightDir, wo);
          if (lensq(k) > 0.0f) {
            emit = true;
            const V3 p0 = vadd(surf.x, vscale(0.001f, lightDir));
            const V3 dl = vsub(lightPos, p0);
            r = makeRay(p0, vnorm<true>(dl));
            busy = travBegin(sc, r, len(dl), s);
            const V3 intensity = vscale(rcp(lensq(pointToLight)), spectrum);
            contrib = vmul(vscale(fabsf(dot(lightDir, surf.n)), k), intensity);
            ++nProbes;
          }
        } else {
          direct[level] = total;
          levels = level + 1;
          alive = level + 1 < depth;
          if (alive) {
            lo = vadd(surf.x, vscale(0.001f, refl)); ld = refl;
            r = makeRay(lo, ld);
            busy = travBegin(sc, r, 1e6f, s);
            ++nSecondary;
          }
        }
      }
    }
    if (valid) {
      V3 acc = mk(0.0f, 0.0f, 0.0f);
      for (int k = levels - 1; k >= 0; --k) acc = vadd(vmul(weight[k], acc), direct[k]);
      storeFinal(W, (size_t)(sLocal * W.framePixels + pixel), acc.x, acc.y, acc.z);
    }
    if (W.rowFlags) rowsSignal(W, valid ? (uint32_t)W.rowOfV[v] : 0u, valid, lane);       // streamed host output
    const uint32_t wp = __reduce_add_sync(kFull, nProbes), ws = __reduce_add_sync(kFull, nSecondary);
    if (lane == 0 && wp) atomicAdd(&W.work[3], wp);
    if (lane == 0 && ws) atomicAdd(&W.work[4], ws);
  }
}

#endif   // !YB_COUNT_WORK

// ---------------------------------------------------------------------------------------------
template <bool ORDERED, bool WIDE, int MIN_BLOCKS, int SH, bool CMP = false>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_wf_shadow(const __grid_constant__ WavefrontParams W) {
  YB_STACK(SH);
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t nEntries = W.dense ? W.nItems * W.samplesPerLaunch * W.base.sc.nSlots : W.work[2];
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&W.work[1], 32u);
    base = __shfl_sync(kFull, base, 0);
    if (base >= nEntries) break;
    const uint32_t entry = base + lane;
    Ray r = makeRay(mk(0, 0, 0), mk(0, 0, 1));
    Trav s;
    s.cur = 0; s.sp = 0; s.tMax = 0.0f; s.best = kNoHit;
    YB_CNT_INIT(s);
    bool busy = false, probe = false;
    if (entry < nEntries) {
      const float4 a = W.q0[entry], b = W.q1[entry];
      if (a.w < 0.0f) {
        W.visibility[entry] = 0;                       // empty dense slot
      } else {
        probe = true;
        r = makeRay(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z));
        busy = travBegin(W.base.sc, r, a.w, s);
      }
    }
    if (WIDE) traverseWarpWide<true, CMP>(W.base.sc, r, s, stack, busy, (int)W.leafThreshold, W.leafRun != 0);
    else traverseWarp<true, ORDERED>(W.base.sc, r, s, stack, busy, (int)W.leafThreshold, W.packed != 0);
    flushCounts(W, s, 1, false);
    uint32_t row = 0;
    if (probe) row = shadowResult(W, entry, s.best == kNoHit);
    if (W.rowFlags) rowsSignal(W, row, probe, lane);
  }
}

// Several lights: sum the visible contributions of a pixel in light order (sum = foldl (+) 0).
__global__ void __launch_bounds__(256) k_wf_resolve(const __grid_constant__ WavefrontParams W) {
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;       // (sample of the launch, item) pair
  if (item >= W.nItems * W.samplesPerLaunch) return;
  const uint32_t nL = W.base.sc.nSlots;
  V3 total = mk(0.0f, 0.0f, 0.0f);
  bool any = false;
  uint32_t pixel = 0;
  for (uint32_t li = 0; li < nL; ++li) {
    const uint32_t e = item * nL + li;
    if (W.visibility[e]) {
      const float4 c = W.q2[e];
      pixel = __float_as_uint(W.q1[e].w);
      total = vadd(total, mk(c.x, c.y, c.z));
      any = true;
    }
  }
  if (!any) return;
  float* out = W.sampleOut + 3 * (size_t)pixel;
  out[0] += total.x; out[1] += total.y; out[2] += total.z;
}

// spp > 1: acc = acc + sample for the samples of this launch in sample order, then pixel = acc / spp after the last.
__global__ void __launch_bounds__(256) k_wf_accum(const __grid_constant__ WavefrontParams W) {
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= W.nItems) return;
  int u, v;
  itemPixel(W, item, u, v);
  const size_t p = 3 * (size_t)(W.base.width * v + u);
  const float n = (float)W.base.spp;
  const bool last = (int)(W.sample + W.samplesPerLaunch) == W.base.spp;
  for (int c = 0; c < 3; ++c) {
    float a = W.sample == 0 ? 0.0f : W.accum[p + c];
    for (uint32_t k = 0; k < W.samplesPerLaunch; ++k) a = a + W.sampleOut[3 * (size_t)k * W.framePixels + p + c];
    if (last) W.base.rgb[p + c] = __fdiv_rn(a, n);
    else W.accum[p + c] = a;
  }
}

#ifndef YB_COUNT_WORK
// item -> pixel table of a tile set (built once, when the tile set is first used)
__global__ void __launch_bounds__(256) k_wf_pixel_table(const __grid_constant__ WavefrontParams W, uint32_t* table) {
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= W.nItems) return;
  int u, v;
  itemPixel(W, item, u, v);
  table[item + W.itemBase] = (uint32_t)u | ((uint32_t)v << 16);
}

cudaError_t launchPixelTable(WavefrontParams W, uint32_t* table, cudaStream_t stream) {
  if (W.nItems == 0) return cudaSuccess;
  W.itemPixels = nullptr;
  k_wf_pixel_table<<<(W.nItems + 255u) / 256u, 256, 0, stream>>>(W, table);
  return cudaGetLastError();
}
#endif

__global__ void k_wf_count(const __grid_constant__ WavefrontParams W) {
  // ray counters for the stats: primary = items, shadow = probes emitted
  atomicAdd(&W.base.counters[0], (unsigned long long)W.nItems * W.samplesPerLaunch);
  atomicAdd(&W.base.counters[1], (unsigned long long)W.work[3]);
  atomicAdd(&W.base.counters[2], (unsigned long long)W.work[4]);
  if (W.bandStat) *W.bandStat += W.work[3];
  W.work[0] = 0; W.work[1] = 0; W.work[2] = 0; W.work[3] = 0; W.work[4] = 0; W.work[5] = 0;   // ready for the next launch
}

// Persistent launch: exactly as many 128-thread CTAs as can be resident (occupancy API), so that every
// CTA is scheduled in the first wave and pulls work until the queue is empty.
static void launchPersistent(void (*kernel)(WavefrontParams), const WavefrontParams& W, int numSMs,
                             cudaStream_t stream, int maxPerSM = 255) {
  // (asking for the smallest shared-memory carve-out instead of the driver's default 32 KB: primary unchanged, shadow
  // kernel 0.426 -> 0.453 ms; not done.  profiles/r1ab section 7)
  int perSM = 0;
  if (W.blocksPerSM) perSM = (int)W.blocksPerSM;
  else if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, 128, 0) != cudaSuccess || perSM < 1) perSM = 8;
  if (perSM > maxPerSM) perSM = maxPerSM;
  kernel<<<numSMs * perSM, 128, 0, stream>>>(W);
}

// kernel<..., SH> for the run-time choice of shared-memory stack entries per lane: 0 (default) or 12 (kept so that the
// negative result of profiles/r2a stays reproducible)
#define YB_PICK_SH(sh, K0, K12) ((sh) >= 8u ? (K12) : (K0))

#ifdef YB_COUNT_WORK
cudaError_t launchWavefrontCounted(WavefrontParams W, int numSMs, cudaStream_t stream, uint32_t* launches,
                                   cudaEvent_t* phaseEvents) {
#else
cudaError_t launchWavefront(WavefrontParams W, int numSMs, cudaStream_t stream, uint32_t* launches,
                            cudaEvent_t* phaseEvents) {
#endif
  if (W.nItems == 0) return cudaSuccess;
  const bool ordered = W.base.traversal == 1;
  const uint32_t itemBlocks = (W.nItems + 255u) / 256u;
  const uint32_t perLaunch = W.samplesPerLaunch ? W.samplesPerLaunch : 1u;
  const uint32_t sh = W.stackShared;
  W.itemsPadded = (W.nItems + 31u) & ~31u;
  W.framePixels = (uint32_t)(W.base.width * W.base.height);
  for (int s = 0; s < W.base.spp; s += (int)perLaunch) {
    W.sample = (uint32_t)s;
    W.samplesPerLaunch = (uint32_t)(W.base.spp - s) < perLaunch ? (uint32_t)(W.base.spp - s) : perLaunch;
    W.sampleOut = W.base.spp == 1 ? W.base.rgb : W.sampleBuf;
    const bool timed = phaseEvents && s == 0;      // phase times of the first launch
    if (timed) cudaEventRecord(phaseEvents[0], stream);
    const bool wide = W.wideTree && !ordered && W.base.sc.wide != nullptr;
    // AREA: the scene has area lights (extension); kept out of the default instantiation
    const bool area = W.base.sc.nAreaLights != 0;
    // the compressed 64-byte nodes (two-kernel set without area lights; the per-batch kernels keep the exact nodes)
    const bool cmp = wide && !area && W.compressed && W.base.sc.cwide != nullptr;
    auto tail = [&]() {
      if (W.base.spp > 1) { k_wf_accum<<<itemBlocks, 256, 0, stream>>>(W); if (launches) *launches += 1; }
      k_wf_count<<<1, 1, 0, stream>>>(W);
      if (launches) *launches += 1;
    };
    auto singleKernelEvents = [&]() {
      if (timed) { cudaEventRecord(phaseEvents[1], stream); cudaEventRecord(phaseEvents[2], stream); cudaEventRecord(phaseEvents[3], stream); }
      if (launches) *launches += 1;
    };
#ifndef YB_COUNT_WORK
    if (W.base.depth != 1) {
      // recursion: the per-batch kernels (planFrame admits them for point lights only, on the 4-wide tree)
      if (W.base.sc.nLights == 1) launchPersistent(k_wf_fused_depth<6>, W, numSMs, stream);
      else launchPersistent(k_wf_fused_depth_lights<6>, W, numSMs, stream);
      singleKernelEvents();
      tail();
      continue;
    }
#endif
    const bool oneSlot = wide && !area && !W.dense && W.base.sc.nSlots <= 1u;
    if (oneSlot && W.persist) {
      launchPersistent(k_wf_persist<8, 0>, W, numSMs, stream, 16);  // the rings in q0 / q2 are sized for 16 CTAs per SM
      singleKernelEvents();
      tail();
      continue;
    }
#ifndef YB_COUNT_WORK
    if (oneSlot && W.fused) {
      if (W.fused & 2u) launchPersistent(k_wf_fused<7>, W, numSMs, stream);     // experiment: 72 registers
      else launchPersistent(k_wf_fused<8>, W, numSMs, stream);
      singleKernelEvents();
      tail();
      continue;
    }
#endif
    if (wide && W.split && W.hits) {
      if (cmp) launchPersistent(k_wf_trace<10, true>, W, numSMs, stream);
      else launchPersistent(k_wf_trace<10, false>, W, numSMs, stream);
      if (timed) cudaEventRecord(phaseEvents[1], stream);
      const uint32_t threads = W.itemsPadded * W.samplesPerLaunch;
      if (area) k_wf_shade<true><<<(threads + 127u) / 128u, 128, 0, stream>>>(W);
      else k_wf_shade<false><<<(threads + 127u) / 128u, 128, 0, stream>>>(W);
      if (launches) *launches += 1;
    } else if (area)
      launchPersistent(wide ? k_wf_primary<false, 8, true, true, 0>
                            : (ordered ? k_wf_primary<true, 8, false, true, 0> : k_wf_primary<false, 8, false, true, 0>),
                       W, numSMs, stream);
    else if (wide && cmp)
      launchPersistent(k_wf_primary<false, 8, true, false, 0, true>, W, numSMs, stream);
    else if (wide)
      launchPersistent(W.capRegisters ? YB_PICK_SH(sh, (k_wf_primary<false, 8, true, false, 0>), (k_wf_primary<false, 8, true, false, 12>))
                                      : k_wf_primary<false, 1, true, false, 0>, W, numSMs, stream);
    else
      launchPersistent(W.capRegisters ? (ordered ? k_wf_primary<true, 8, false, false, 0> : k_wf_primary<false, 8, false, false, 0>)
                                      : (ordered ? k_wf_primary<true, 1, false, false, 0> : k_wf_primary<false, 1, false, false, 0>),
                       W, numSMs, stream);
    if (timed && !(wide && W.split && W.hits)) cudaEventRecord(phaseEvents[1], stream);
    if (timed) cudaEventRecord(phaseEvents[2], stream);
    // 10 CTAs / SM (48 registers) measured 2-3 % faster than the unconstrained 56 registers / 9 CTAs
    if (wide && cmp) launchPersistent(k_wf_shadow<false, true, 10, 0, true>, W, numSMs, stream);
    else if (wide) launchPersistent(YB_PICK_SH(sh, (k_wf_shadow<false, true, 10, 0>), (k_wf_shadow<false, true, 10, 12>)), W, numSMs, stream);
    else launchPersistent(ordered ? k_wf_shadow<true, false, 1, 0> : k_wf_shadow<false, false, 1, 0>, W, numSMs, stream);
    if (timed) cudaEventRecord(phaseEvents[3], stream);
    if (launches) *launches += 2;
    if (W.dense) {
      k_wf_resolve<<<(W.nItems * W.samplesPerLaunch + 255u) / 256u, 256, 0, stream>>>(W);
      if (launches) *launches += 1;
    }
    tail();
  }
  return cudaGetLastError();
}

YB_WF_CLOSE
