// bvh_build_gpu.cu -- construction of the reference's BVH ON THE GPU.
//
// Same tree as bvh_build.cpp / Culling.hs:27-60 (Midpoint split): same split dimension (maxExtent of the
// centroid bounds, Vectors.hs:62-66), same midpoint 0.5*cmin + 0.5*cmax, same stable partition
// `centroid <= midpoint`, same depth cap and multi-leaf rules, hence the same left-first DFS primitive
// order (= tie-break rank) and the same boxes (min / max are exact, so a node's box is the min / max
// over its primitives in any order).  All arithmetic is binary32 without FMA contraction.
//
//   phase 0  bounds + centroids + validation per primitive                       (k_prim_bounds)
//   phase 1  level-synchronous splitting of the LARGE segments: centroid / primitive bounds by
//            warp-aggregated atomics, one global exclusive scan per level for the stable partition
//   phase 2  one thread per SMALL segment (<= kSmall primitives) finishes its subtree serially
//   phase 3  pre-order numbering of the inner nodes = sort by (first position asc, last position desc),
//            emission of the 64-byte traversal nodes, permutation of the primitives into DFS order
//
// SurfaceAreaHeuristic (Culling.hs:62-112): the 16 bucket counts / boxes of every segment are accumulated with
// atomics (min / max are exact), then ONE thread per segment evaluates the reference's cost expression in the
// reference's order (wrappedArea = left fold over the buckets, prefix / suffix joins, minimumBy with its NaN
// behaviour) -- same floats, same split.  Segments of at most 4 primitives split at the midpoint (Culling.hs:110-112).
// The degenerate case of an empty LEFT partition is left to the host builder (bvh_build.cpp).
#include "bvh_build_gpu.hpp"

#include <cub/cub.cuh>

namespace yb {
namespace {

constexpr uint32_t kSmall = 64;            // segments up to this size are finished by one thread
constexpr uint32_t kInvalid = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t encF(float f) {      // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float decF(uint32_t e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}
// centroid = 0.5 @* pMin + 0.5 @* pMax   (AABBs.hs:48-49)
__device__ __forceinline__ float centroid1(float lo, float hi) {
  return __fadd_rn(__fmul_rn(0.5f, lo), __fmul_rn(0.5f, hi));
}
// maxDimension (Vectors.hs:62-66)
__device__ __forceinline__ int maxDim(float x, float y, float z) { return (x > y && x > z) ? 0 : (y > z ? 1 : 2); }
__device__ __forceinline__ float comp(const float4& v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }

// ---- SurfaceAreaHeuristic helpers (Culling.hs:62-112; host twin: bvh_build.cpp sahSplit) ------------------
constexpr int kBuckets = 16;
// bucketId = clamp (floor (16 * (c - cmin) / (cmax - cmin))); `floor x :: Int` of NaN / out of range is INT64_MIN
__device__ __forceinline__ int bucketOf(float c, float centMin, float extent) {
  const float frac = __fdiv_rn(__fsub_rn(c, centMin), extent);
  const float f = floorf(__fmul_rn((float)kBuckets, frac));
  if (!(f >= -9.2233720368547758e18f && f < 9.2233720368547758e18f)) return 0;     // INT64_MIN clamps to 0
  if (f < 0.0f) return 0;
  if (f > (float)(kBuckets - 1)) return kBuckets - 1;
  return (int)f;
}
struct BoxD { float lo[3], hi[3]; };
__device__ __forceinline__ BoxD emptyBoxD() {
  BoxD b;
  for (int c = 0; c < 3; ++c) { b.lo[c] = INFINITY; b.hi[c] = -INFINITY; }
  return b;
}
// GHC class-default min / max (select semantics); inputs are never NaN here
__device__ __forceinline__ BoxD joinD(const BoxD& a, const BoxD& b) {
  BoxD r;
  for (int c = 0; c < 3; ++c) { r.lo[c] = a.lo[c] <= b.lo[c] ? a.lo[c] : b.lo[c]; r.hi[c] = a.hi[c] <= b.hi[c] ? b.hi[c] : a.hi[c]; }
  return r;
}
__device__ __forceinline__ float surfD(const BoxD& b) {                    // AABBs.hs:51-53
  const float dx = __fsub_rn(b.hi[0], b.lo[0]), dy = __fsub_rn(b.hi[1], b.lo[1]), dz = __fsub_rn(b.hi[2], b.lo[2]);
  return __fmul_rn(2.0f, __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dx, dz)), __fmul_rn(dy, dz)));
}
// The split bucket: primitives of buckets <= result go left.
__device__ int sahBestSplit(const int* counts, const BoxD* boxes) {
  float wrappedArea = 0.0f;
  for (int b = 0; b < kBuckets; ++b) wrappedArea = __fadd_rn(wrappedArea, surfD(boxes[b]));
  int bestSplit = 0;
  float bestCost = 0.0f;
  BoxD pre = emptyBoxD();
  int cpre = 0;
  for (int s = 0; s <= kBuckets - 2; ++s) {
    pre = joinD(pre, boxes[s]);
    cpre += counts[s];
    BoxD suf = emptyBoxD();
    int csuf = 0;
    for (int b = s + 1; b < kBuckets; ++b) { suf = joinD(suf, boxes[b]); csuf += counts[b]; }
    const float cost = __fadd_rn(0.125f, __fdiv_rn(__fadd_rn(__fmul_rn((float)cpre, surfD(pre)), __fmul_rn((float)csuf, surfD(suf))),
                                                   wrappedArea));
    // minimumBy (compare `on` snd): the accumulator is replaced unless acc < next or acc == next
    if (s == 0 || !((bestCost < cost) || (bestCost == cost))) { bestSplit = s; bestCost = cost; }
  }
  return bestSplit;
}

// ---- phase 0 ------------------------------------------------------------------------------------------
__global__ void k_prim_bounds(GpuBuildInput in, float4* blo, float4* bhi, float4* cen, uint32_t* idx, uint32_t* segOf,
                              uint32_t* errFlags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in.nPrims) return;
  const uint32_t ref = in.primOrder ? in.primOrder[i] : (i < in.nSpheres ? i : (0x80000000u | (i - in.nSpheres)));
  const uint32_t k = ref & 0x7FFFFFFFu;
  float lo[3], hi[3];
  uint32_t mat;
  bool finite = true;
  if (ref >> 31) {
    if (k >= in.nTriangles) { atomicOr(errFlags, 4u); return; }
    // boundTriangle = includePoint (fromPoints p0 p1) p2   (Shapes.hs:62-63)
    for (int c = 0; c < 3; ++c) {
      const float a = in.triP0[3 * (size_t)k + c], b = in.triP1[3 * (size_t)k + c], d = in.triP2[3 * (size_t)k + c];
      finite = finite && isfinite(a) && isfinite(b) && isfinite(d);
      lo[c] = fminf(fminf(a, b), d);
      hi[c] = fmaxf(fmaxf(a, b), d);
    }
    mat = in.triMaterial ? in.triMaterial[k] : 0u;
  } else {
    if (k >= in.nSpheres) { atomicOr(errFlags, 4u); return; }
    // boundSphere r s = fromPoints (s + vof r) (s - vof r)  (Shapes.hs:32-33)
    const float r = in.sphRadius[k];
    finite = isfinite(r);
    for (int c = 0; c < 3; ++c) {
      const float s = in.sphCenter[3 * (size_t)k + c];
      finite = finite && isfinite(s);
      const float x = __fadd_rn(s, r), y = __fsub_rn(s, r);
      lo[c] = fminf(x, y);
      hi[c] = fmaxf(x, y);
    }
    mat = in.sphMaterial ? in.sphMaterial[k] : 0u;
  }
  if (!finite) atomicOr(errFlags, 1u);
  if (mat >= in.nMaterials) atomicOr(errFlags, 2u);
  blo[i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
  bhi[i] = make_float4(hi[0], hi[1], hi[2], 0.0f);
  cen[i] = make_float4(centroid1(lo[0], hi[0]), centroid1(lo[1], hi[1]), centroid1(lo[2], hi[2]), 0.0f);
  idx[i] = i;
  segOf[i] = 0;
}

// ---- phase 1 ------------------------------------------------------------------------------------------
struct Segs {            // active (large) segments of the current level
  uint32_t* lo; uint32_t* hi; uint32_t* node;
  uint32_t* acc;         // 12 ordered-uint accumulators per segment: cmin3 cmax3 bmin3 bmax3
  uint32_t* mode;        // 0 = became a leaf, 1 = split
  uint32_t* dim; float* mid;
  float* cmin; float* extent;          // SAH: centroid bounds along `dim`
  uint32_t* split;                     // SAH: split bucket, or kInvalid = midpoint rule (<= 4 primitives never happens here)
  uint32_t* bacc;                      // SAH: per segment 16 x (count, lo3, hi3) ordered-uint accumulators
  uint32_t* leftCount; uint32_t* leftSeg; uint32_t* rightSeg;   // child segment ids (kInvalid: small / leaf)
};

__global__ void k_acc_init(uint32_t* acc, uint32_t nSeg) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSeg * 12u) return;
  const uint32_t k = i % 12u;
  acc[i] = (k < 3u || (k >= 6u && k < 9u)) ? 0xFFFFFFFFu : 0u;     // mins start high, maxes low
}

__global__ void k_seg_bounds(uint32_t n, const uint32_t* idx, const uint32_t* segOf, const float4* blo, const float4* bhi,
                             const float4* cen, uint32_t* acc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t seg = i < n ? segOf[i] : kInvalid;
  uint32_t v[12];
  if (seg != kInvalid) {
    const uint32_t p = idx[i];
    const float4 c = cen[p], l = blo[p], h = bhi[p];
    v[0] = encF(c.x); v[1] = encF(c.y); v[2] = encF(c.z); v[3] = v[0]; v[4] = v[1]; v[5] = v[2];
    v[6] = encF(l.x); v[7] = encF(l.y); v[8] = encF(l.z); v[9] = encF(h.x); v[10] = encF(h.y); v[11] = encF(h.z);
  }
  const unsigned full = 0xFFFFFFFFu;
  const uint32_t seg0 = __shfl_sync(full, seg, 0);
  if (__all_sync(full, seg == seg0)) {                 // whole warp inside one segment: one atomic per value
    if (seg0 == kInvalid) return;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const bool isMin = k < 3 || (k >= 6 && k < 9);
      const uint32_t r = isMin ? __reduce_min_sync(full, v[k]) : __reduce_max_sync(full, v[k]);
      if ((threadIdx.x & 31) == 0) { if (isMin) atomicMin(&acc[seg0 * 12u + k], r); else atomicMax(&acc[seg0 * 12u + k], r); }
    }
  } else if (seg != kInvalid) {
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const bool isMin = k < 3 || (k >= 6 && k < 9);
      if (isMin) atomicMin(&acc[seg * 12u + k], v[k]); else atomicMax(&acc[seg * 12u + k], v[k]);
    }
  }
}

__device__ __forceinline__ void writeBox(GpuNode& nd, const uint32_t* a) {
  nd.box[0] = decF(a[6]); nd.box[1] = decF(a[7]); nd.box[2] = decF(a[8]);
  nd.box[3] = decF(a[9]); nd.box[4] = decF(a[10]); nd.box[5] = decF(a[11]);
}

// buildTree (Culling.hs:31-50) decision for every large segment of this level.
__global__ void k_seg_decide(Segs s, uint32_t nSeg, int depthLeft, uint32_t level, GpuNode* nodes) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSeg) return;
  const uint32_t* a = s.acc + 12u * i;
  GpuNode& nd = nodes[s.node[i]];
  nd.lo = s.lo[i]; nd.hi = s.hi[i]; nd.left = -1; nd.right = -1; nd.level = level;
  writeBox(nd, a);
  s.leftSeg[i] = kInvalid; s.rightSeg[i] = kInvalid; s.leftCount[i] = 0;
  if (depthLeft == 0) { nd.kind = 3; s.mode[i] = 0; return; }           // buildTree 0 xs = multiLeaf xs
  const float cminx = decF(a[0]), cminy = decF(a[1]), cminz = decF(a[2]);
  const float cmaxx = decF(a[3]), cmaxy = decF(a[4]), cmaxz = decF(a[5]);
  const int d = maxDim(__fsub_rn(cmaxx, cminx), __fsub_rn(cmaxy, cminy), __fsub_rn(cmaxz, cminz));
  const float lo = d == 0 ? cminx : (d == 1 ? cminy : cminz), hi = d == 0 ? cmaxx : (d == 1 ? cmaxy : cmaxz);
  s.dim[i] = (uint32_t)d;
  s.mid[i] = centroid1(lo, hi);                                         // getDimension dim (centroid bbOfCentroids)
  s.cmin[i] = lo;
  s.extent[i] = __fsub_rn(hi, lo);
  s.split[i] = kInvalid;
  s.mode[i] = 1;
}

// SAH: bucket counts and boxes of every splitting segment
__global__ void k_bucket_init(uint32_t* bacc, uint32_t nSeg) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSeg * kBuckets * 7u) return;
  const uint32_t k = i % 7u;
  bacc[i] = k == 0u ? 0u : (k <= 3u ? 0xFFFFFFFFu : 0u);               // count, mins high, maxes low
}

__global__ void k_bucket_accum(uint32_t n, const uint32_t* idx, const uint32_t* segOf, const float4* blo, const float4* bhi,
                               const float4* cen, Segs s) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t seg = segOf[i];
  if (seg == kInvalid || !s.mode[seg]) return;
  const uint32_t p = idx[i];
  const int b = bucketOf(comp(cen[p], (int)s.dim[seg]), s.cmin[seg], s.extent[seg]);
  uint32_t* a = s.bacc + ((size_t)seg * kBuckets + b) * 7u;
  const float4 l = blo[p], h = bhi[p];
  atomicAdd(&a[0], 1u);
  atomicMin(&a[1], encF(l.x)); atomicMin(&a[2], encF(l.y)); atomicMin(&a[3], encF(l.z));
  atomicMax(&a[4], encF(h.x)); atomicMax(&a[5], encF(h.y)); atomicMax(&a[6], encF(h.z));
}

__global__ void k_sah_decide(Segs s, uint32_t nSeg) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSeg || !s.mode[i]) return;
  int counts[kBuckets];
  BoxD boxes[kBuckets];
  for (int b = 0; b < kBuckets; ++b) {
    const uint32_t* a = s.bacc + ((size_t)i * kBuckets + b) * 7u;
    counts[b] = (int)a[0];
    if (a[0] == 0u) boxes[b] = emptyBoxD();
    else for (int c = 0; c < 3; ++c) { boxes[b].lo[c] = decF(a[1 + c]); boxes[b].hi[c] = decF(a[4 + c]); }
  }
  s.split[i] = (uint32_t)sahBestSplit(counts, boxes);
}

__global__ void k_flags(uint32_t n, const uint32_t* idx, const uint32_t* segOf, const float4* cen, Segs s,
                        uint32_t* flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  uint32_t f = 0;
  if (i < n) {
    const uint32_t seg = segOf[i];
    if (seg != kInvalid && s.mode[seg]) {
      const float c = comp(cen[idx[i]], (int)s.dim[seg]);
      if (s.split[seg] == kInvalid) f = c <= s.mid[seg] ? 1u : 0u;                               // midpointSplit
      else f = (uint32_t)bucketOf(c, s.cmin[seg], s.extent[seg]) <= s.split[seg] ? 1u : 0u;       // sahSplit
    }
  }
  flags[i] = f;
}

// Creates the children of every split segment.  counters: [0] nodes, [1] next-level segments,
// [2] small segments, [3] unsupported flag, [4] max level.
__global__ void k_seg_children(Segs s, uint32_t nSeg, const uint32_t* scan, GpuNode* nodes, uint32_t* counters,
                               uint32_t* nextLo, uint32_t* nextHi, uint32_t* nextNode, uint32_t* smallLo, uint32_t* smallHi,
                               uint32_t* smallNode) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSeg || !s.mode[i]) return;
  const uint32_t lo = s.lo[i], hi = s.hi[i], count = hi - lo;
  const uint32_t leftCount = scan[hi] - scan[lo];
  GpuNode& nd = nodes[s.node[i]];
  s.leftCount[i] = leftCount;
  if (leftCount == count) { nd.kind = 3; s.mode[i] = 0; return; }        // if null xs2 then multiLeaf xs1
  if (leftCount == 0) { atomicExch(&counters[3], 1u); nd.kind = 3; s.mode[i] = 0; return; }   // host builder handles it
  const uint32_t base = atomicAdd(&counters[0], 2u);
  nd.kind = 2; nd.left = (int32_t)base; nd.right = (int32_t)(base + 1);
  const uint32_t mid = lo + leftCount;
  const uint32_t clo[2] = {lo, mid}, chi[2] = {mid, hi};
  for (int c = 0; c < 2; ++c) {
    const uint32_t n = chi[c] - clo[c];
    uint32_t segId = kInvalid;
    if (n > kSmall) {
      segId = atomicAdd(&counters[1], 1u);
      nextLo[segId] = clo[c]; nextHi[segId] = chi[c]; nextNode[segId] = base + c;
    } else {
      const uint32_t k = atomicAdd(&counters[2], 1u);
      smallLo[k] = clo[c]; smallHi[k] = chi[c]; smallNode[k] = base + c;
      nodes[base + c].level = nd.level + 1;
    }
    if (c == 0) s.leftSeg[i] = segId; else s.rightSeg[i] = segId;
  }
}

__global__ void k_scatter(uint32_t n, const uint32_t* idxIn, uint32_t* idxOut, const uint32_t* segOf, uint32_t* segOfNext,
                          const uint32_t* scan, Segs s) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t seg = segOf[i];
  if (seg == kInvalid || !s.mode[seg]) { idxOut[i] = idxIn[i]; segOfNext[i] = kInvalid; return; }
  const uint32_t lo = s.lo[seg];
  const uint32_t before = scan[i] - scan[lo];                  // left-going elements before i in the segment
  const bool left = scan[i + 1] != scan[i];
  const uint32_t pos = left ? lo + before : lo + s.leftCount[seg] + ((i - lo) - before);   // stable partition
  idxOut[pos] = idxIn[i];
  segOfNext[pos] = left ? s.leftSeg[seg] : s.rightSeg[seg];
}

// ---- phase 2 ------------------------------------------------------------------------------------------
// One thread finishes the subtree of a segment of at most kSmall primitives (iterative buildTree).
__global__ void k_finish_small(uint32_t nSmall, const uint32_t* smallLo, const uint32_t* smallHi, const uint32_t* smallNode,
                               uint32_t* idx, const float4* blo, const float4* bhi, const float4* cen, GpuNode* nodes,
                               uint32_t* counters, int maxDepth, int splitMode) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nSmall) return;
  uint32_t stLo[kSmall + 1], stHi[kSmall + 1], stNode[kSmall + 1];
  uint32_t tmp[kSmall];
  int sp = 0;
  stLo[0] = smallLo[t]; stHi[0] = smallHi[t]; stNode[0] = smallNode[t];
  sp = 1;
  uint32_t maxLevel = 0;
  while (sp > 0) {
    --sp;
    const uint32_t lo = stLo[sp], hi = stHi[sp], id = stNode[sp], count = hi - lo;
    GpuNode& nd = nodes[id];
    const uint32_t level = nd.level;
    maxLevel = max(maxLevel, level);
    nd.lo = lo; nd.hi = hi; nd.left = -1; nd.right = -1;
    float bl[3] = {INFINITY, INFINITY, INFINITY}, bh[3] = {-INFINITY, -INFINITY, -INFINITY};
    float cl[3] = {INFINITY, INFINITY, INFINITY}, ch[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t k = lo; k < hi; ++k) {
      const uint32_t p = idx[k];
      const float4 l = blo[p], h = bhi[p], c = cen[p];
      bl[0] = fminf(bl[0], l.x); bl[1] = fminf(bl[1], l.y); bl[2] = fminf(bl[2], l.z);
      bh[0] = fmaxf(bh[0], h.x); bh[1] = fmaxf(bh[1], h.y); bh[2] = fmaxf(bh[2], h.z);
      cl[0] = fminf(cl[0], c.x); cl[1] = fminf(cl[1], c.y); cl[2] = fminf(cl[2], c.z);
      ch[0] = fmaxf(ch[0], c.x); ch[1] = fmaxf(ch[1], c.y); ch[2] = fmaxf(ch[2], c.z);
    }
    for (int c = 0; c < 3; ++c) { nd.box[c] = bl[c]; nd.box[3 + c] = bh[c]; }
    if (count == 1) { nd.kind = 1; continue; }                                 // single-primitive leaf
    if (maxDepth - (int)level == 0) { nd.kind = 3; continue; }                 // depth cap: multi-leaf
    const int d = maxDim(__fsub_rn(ch[0], cl[0]), __fsub_rn(ch[1], cl[1]), __fsub_rn(ch[2], cl[2]));
    const float mid = centroid1(cl[d], ch[d]);
    int sahSplit = -1;                                                          // -1: midpoint rule
    const float centMin = cl[d], extent = __fsub_rn(ch[d], cl[d]);
    if (splitMode == 1 && count > 4u) {                                         // sahSplit (Culling.hs:62-112)
      int counts[kBuckets];
      BoxD boxes[kBuckets];
      for (int b = 0; b < kBuckets; ++b) { counts[b] = 0; boxes[b] = emptyBoxD(); }
      for (uint32_t k = lo; k < hi; ++k) {
        const uint32_t p = idx[k];
        const int b = bucketOf(comp(cen[p], d), centMin, extent);
        const float4 l4 = blo[p], h4 = bhi[p];
        BoxD pb;
        pb.lo[0] = l4.x; pb.lo[1] = l4.y; pb.lo[2] = l4.z; pb.hi[0] = h4.x; pb.hi[1] = h4.y; pb.hi[2] = h4.z;
        counts[b]++;
        boxes[b] = joinD(boxes[b], pb);
      }
      sahSplit = sahBestSplit(counts, boxes);
    }
    uint32_t l = lo, r = 0;
    for (uint32_t k = lo; k < hi; ++k) {                                        // stable partition
      const uint32_t p = idx[k];
      const float c = comp(cen[p], d);
      const bool left = sahSplit < 0 ? (c <= mid) : (bucketOf(c, centMin, extent) <= sahSplit);
      if (left) idx[l++] = p; else tmp[r++] = p;
    }
    for (uint32_t k = 0; k < r; ++k) idx[l + k] = tmp[k];
    if (l == hi) { nd.kind = 3; continue; }                                    // right side empty: multi-leaf
    if (l == lo) { atomicExch(&counters[3], 1u); nd.kind = 3; continue; }      // left side empty: host builder
    const uint32_t base = atomicAdd(&counters[0], 2u);
    nd.kind = 2; nd.left = (int32_t)base; nd.right = (int32_t)(base + 1);
    nodes[base].level = level + 1; nodes[base + 1].level = level + 1;
    // push right first so that the left child is processed next (order is irrelevant for the result)
    stLo[sp] = l; stHi[sp] = hi; stNode[sp] = base + 1; ++sp;
    stLo[sp] = lo; stHi[sp] = l; stNode[sp] = base; ++sp;
  }
  atomicMax(&counters[4], maxLevel);
}

// ---- phase 3 ------------------------------------------------------------------------------------------
__global__ void k_node_keys(uint32_t nNodes, uint32_t nPrims, const GpuNode* nodes, unsigned long long* keys, uint32_t* ids,
                            uint32_t* counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nNodes) return;
  const GpuNode& nd = nodes[i];
  ids[i] = i;
  if (nd.kind == 2) {
    // pre-order: earlier first position first; among nodes starting at the same position the ancestor
    // (larger last position) first
    keys[i] = ((unsigned long long)nd.lo << 32) | (unsigned long long)(nPrims - nd.hi);
    atomicAdd(&counters[5], 1u);
  } else {
    keys[i] = ~0ull;
    if (nd.kind == 3) atomicAdd(&counters[6], 1u);
  }
}

__global__ void k_pre_index(uint32_t nInner, const uint32_t* sortedIds, uint32_t* preIdx) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nInner) preIdx[sortedIds[r]] = r;
}

__device__ __forceinline__ uint32_t childRef(const GpuNode* nodes, int32_t id, const uint32_t* preIdx, uint2* multi,
                                             uint32_t* counters) {
  const GpuNode& c = nodes[id];
  if (c.kind == 2) return preIdx[id];
  if (c.kind == 1) return 0x80000000u | c.lo;
  if (c.kind == 3) {
    const uint32_t k = atomicAdd(&counters[7], 1u);
    multi[k] = make_uint2(c.lo, c.hi - c.lo);
    return 0xC0000000u | k;
  }
  return kInvalid;
}

__global__ void k_emit_flat(uint32_t nInner, const uint32_t* sortedIds, const GpuNode* nodes, const uint32_t* preIdx,
                            float4* flat, uint2* multi, uint32_t* counters) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nInner) return;
  const GpuNode& nd = nodes[sortedIds[r]];
  const GpuNode& L = nodes[nd.left];
  const GpuNode& R = nodes[nd.right];
  flat[4 * (size_t)r + 0] = make_float4(L.box[0], L.box[1], L.box[3], L.box[4]);
  flat[4 * (size_t)r + 1] = make_float4(R.box[0], R.box[1], R.box[3], R.box[4]);
  flat[4 * (size_t)r + 2] = make_float4(L.box[2], L.box[5], R.box[2], R.box[5]);
  const uint32_t refL = childRef(nodes, nd.left, preIdx, multi, counters);
  const uint32_t refR = childRef(nodes, nd.right, preIdx, multi, counters);
  flat[4 * (size_t)r + 3] = make_float4(__uint_as_float(refL), __uint_as_float(refR), 0.0f, 0.0f);
}

// Primitive records and normals in DFS order (same packing as the host path in capi.cu).
__global__ void k_pack_prims(GpuBuildInput in, const uint32_t* idx, float4* recs, float4* nrm) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in.nPrims) return;
  const uint32_t id = idx[i];
  const uint32_t ref = in.primOrder ? in.primOrder[id] : (id < in.nSpheres ? id : (0x80000000u | (id - in.nSpheres)));
  const size_t k = ref & 0x7FFFFFFFu;
  if (ref >> 31) {
    const float* a = in.triP0 + 3 * k; const float* b = in.triP1 + 3 * k; const float* c = in.triP2 + 3 * k;
    const uint32_t mat = in.triMaterial ? in.triMaterial[k] : 0u;
    recs[3 * (size_t)i + 0] = make_float4(a[0], a[1], a[2], __uint_as_float((mat << 1) | 1u));
    recs[3 * (size_t)i + 1] = make_float4(__fsub_rn(b[0], a[0]), __fsub_rn(b[1], a[1]), __fsub_rn(b[2], a[2]), __uint_as_float(id));
    recs[3 * (size_t)i + 2] = make_float4(__fsub_rn(c[0], a[0]), __fsub_rn(c[1], a[1]), __fsub_rn(c[2], a[2]), 0.0f);
    const float* n0 = in.triN0 + 3 * k; const float* n1 = in.triN1 + 3 * k; const float* n2 = in.triN2 + 3 * k;
    nrm[3 * (size_t)i + 0] = make_float4(n0[0], n0[1], n0[2], 0.0f);
    nrm[3 * (size_t)i + 1] = make_float4(n1[0], n1[1], n1[2], 0.0f);
    nrm[3 * (size_t)i + 2] = make_float4(n2[0], n2[1], n2[2], 0.0f);
  } else {
    const float* s = in.sphCenter + 3 * k;
    const uint32_t mat = in.sphMaterial ? in.sphMaterial[k] : 0u;
    recs[3 * (size_t)i + 0] = make_float4(s[0], s[1], s[2], __uint_as_float(mat << 1));
    recs[3 * (size_t)i + 1] = make_float4(in.sphRadius[k], 0.0f, 0.0f, __uint_as_float(id));
    recs[3 * (size_t)i + 2] = make_float4(0, 0, 0, 0);
    nrm[3 * (size_t)i + 0] = nrm[3 * (size_t)i + 1] = nrm[3 * (size_t)i + 2] = make_float4(0, 0, 0, 0);
  }
}

// One stream-ordered arena for every temporary of a build: a single cudaMallocAsync from a pool that
// keeps its memory between builds (dozens of separate cudaMalloc / cudaFree calls used to cost far more
// than the build itself).
struct Arena {
  char* base = nullptr;
  size_t off = 0;
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += (n ? n : 1) * sizeof(T);
    return p;
  }
};

struct BuildBuffers {
  float4 *blo, *bhi, *cen;
  uint32_t *idxA, *idxB, *segA, *segB, *flags, *scan, *counters;
  GpuNode* nodes;
  uint32_t *sLo[2], *sHi[2], *sNode[2], *acc, *mode, *dim, *leftCount, *leftSeg, *rightSeg, *smLo, *smHi, *smNode;
  uint32_t *split, *bacc;
  float *mid, *cmin, *extent;
  unsigned char *scanTemp, *sortTemp;
  unsigned long long *keysIn, *keysOut;
  uint32_t *idsIn, *idsOut, *preIdx;
  void layout(Arena& a, uint32_t n, size_t maxSeg, size_t scanBytes, size_t sortBytes) {
    blo = a.take<float4>(n); bhi = a.take<float4>(n); cen = a.take<float4>(n);
    idxA = a.take<uint32_t>(n); idxB = a.take<uint32_t>(n); segA = a.take<uint32_t>(n); segB = a.take<uint32_t>(n);
    flags = a.take<uint32_t>((size_t)n + 1); scan = a.take<uint32_t>((size_t)n + 1); counters = a.take<uint32_t>(16);
    nodes = a.take<GpuNode>(2 * (size_t)n);
    for (int k = 0; k < 2; ++k) { sLo[k] = a.take<uint32_t>(maxSeg); sHi[k] = a.take<uint32_t>(maxSeg); sNode[k] = a.take<uint32_t>(maxSeg); }
    acc = a.take<uint32_t>(12 * maxSeg); mode = a.take<uint32_t>(maxSeg); dim = a.take<uint32_t>(maxSeg);
    mid = a.take<float>(maxSeg); cmin = a.take<float>(maxSeg); extent = a.take<float>(maxSeg);
    split = a.take<uint32_t>(maxSeg); bacc = a.take<uint32_t>(maxSeg * kBuckets * 7);
    leftCount = a.take<uint32_t>(maxSeg); leftSeg = a.take<uint32_t>(maxSeg); rightSeg = a.take<uint32_t>(maxSeg);
    smLo = a.take<uint32_t>(n); smHi = a.take<uint32_t>(n); smNode = a.take<uint32_t>(n);
    scanTemp = a.take<unsigned char>(scanBytes); sortTemp = a.take<unsigned char>(sortBytes);
    keysIn = a.take<unsigned long long>(2 * (size_t)n); keysOut = a.take<unsigned long long>(2 * (size_t)n);
    idsIn = a.take<uint32_t>(2 * (size_t)n); idsOut = a.take<uint32_t>(2 * (size_t)n); preIdx = a.take<uint32_t>(2 * (size_t)n);
  }
};

void keepPoolMemory() {
  static bool done = false;
  if (done) return;
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  done = true;
}

#define GB(call)                                                         \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) { out.error = e__; out.where = #call; if (arenaBase) cudaFreeAsync(arenaBase, st); return false; } \
  } while (0)

inline unsigned blocks(size_t n, unsigned per = 256) { return (unsigned)((n + per - 1) / per); }

}  // namespace

bool gpuBuildSupported(int splitMode) { return splitMode == 0 || splitMode == 1; }

bool buildBvhOnDevice(const GpuBuildInput& in, int maxDepth, int splitMode, GpuBuildOutput& out) {
  out = GpuBuildOutput();
  const uint32_t n = in.nPrims;
  if (n == 0) return true;                                   // bvh _ _ [] = const Nothing
  cudaStream_t st = nullptr;
  char* arenaBase = nullptr;
  keepPoolMemory();
  // segment arrays: at most n / kSmall large segments per level (+2), small segments at most n
  const size_t maxSeg = n / kSmall + 2;
  size_t scanBytes = 0, sortBytes = 0;
  GB(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n + 1), st));
  GB(cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                     (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(2 * (size_t)n), 0, 64, st));
  BuildBuffers B;
  Arena sizing;
  B.layout(sizing, n, maxSeg, scanBytes, sortBytes);
  GB(cudaMallocAsync((void**)&arenaBase, sizing.off + 256, st));
  Arena arena;
  arena.base = arenaBase;
  B.layout(arena, n, maxSeg, scanBytes, sortBytes);
  GB(cudaMemsetAsync(B.counters, 0, 16 * sizeof(uint32_t), st));

  k_prim_bounds<<<blocks(n), 256, 0, st>>>(in, B.blo, B.bhi, B.cen, B.idxA, B.segA, B.counters + 8);

  // root = node 0
  uint32_t hostCounters[16] = {0};
  hostCounters[0] = 1;                                       // node count
  uint32_t nSeg = 0, cur = 0;
  {
    const uint32_t z = 0, nn = n;
    uint32_t* lo = n > kSmall ? B.sLo[0] : B.smLo;
    uint32_t* hi = n > kSmall ? B.sHi[0] : B.smHi;
    uint32_t* nd = n > kSmall ? B.sNode[0] : B.smNode;
    GB(cudaMemcpyAsync(lo, &z, 4, cudaMemcpyHostToDevice, st));
    GB(cudaMemcpyAsync(hi, &nn, 4, cudaMemcpyHostToDevice, st));
    GB(cudaMemcpyAsync(nd, &z, 4, cudaMemcpyHostToDevice, st));
    if (n > kSmall) nSeg = 1;
    else { hostCounters[2] = 1; GB(cudaMemsetAsync(B.nodes, 0, sizeof(GpuNode), st)); }   // level 0
  }
  GB(cudaMemcpyAsync(B.counters, hostCounters, 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));

  uint32_t* idxIn = B.idxA; uint32_t* idxOut = B.idxB;
  uint32_t* segIn = B.segA; uint32_t* segOut = B.segB;
  uint32_t level = 0;
  while (nSeg > 0) {
    if (level > 200) { out.tooDeep = true; break; }
    Segs s{B.sLo[cur], B.sHi[cur], B.sNode[cur], B.acc,  B.mode,      B.dim,     B.mid,     B.cmin, B.extent,
           B.split,    B.bacc,     B.leftCount,  B.leftSeg, B.rightSeg};
    k_acc_init<<<blocks(12 * (size_t)nSeg), 256, 0, st>>>(B.acc, nSeg);
    k_seg_bounds<<<blocks(n), 256, 0, st>>>(n, idxIn, segIn, B.blo, B.bhi, B.cen, B.acc);
    k_seg_decide<<<blocks(nSeg), 256, 0, st>>>(s, nSeg, maxDepth - (int)level, level, B.nodes);
    if (splitMode == 1) {            // large segments always hold more than 4 primitives: sahSplit proper
      k_bucket_init<<<blocks((size_t)nSeg * kBuckets * 7), 256, 0, st>>>(B.bacc, nSeg);
      k_bucket_accum<<<blocks(n), 256, 0, st>>>(n, idxIn, segIn, B.blo, B.bhi, B.cen, s);
      k_sah_decide<<<blocks(nSeg, 64), 64, 0, st>>>(s, nSeg);
    }
    k_flags<<<blocks((size_t)n + 1), 256, 0, st>>>(n, idxIn, segIn, B.cen, s, B.flags);
    GB(cub::DeviceScan::ExclusiveSum(B.scanTemp, scanBytes, B.flags, B.scan, (int)(n + 1), st));
    GB(cudaMemsetAsync(B.counters + 1, 0, sizeof(uint32_t), st));
    k_seg_children<<<blocks(nSeg), 256, 0, st>>>(s, nSeg, B.scan, B.nodes, B.counters, B.sLo[cur ^ 1], B.sHi[cur ^ 1],
                                                 B.sNode[cur ^ 1], B.smLo, B.smHi, B.smNode);
    k_scatter<<<blocks(n), 256, 0, st>>>(n, idxIn, idxOut, segIn, segOut, B.scan, s);
    GB(cudaMemcpyAsync(hostCounters, B.counters, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GB(cudaStreamSynchronize(st));
    nSeg = hostCounters[1];
    cur ^= 1;
    uint32_t* t1 = idxIn; idxIn = idxOut; idxOut = t1;
    uint32_t* t2 = segIn; segIn = segOut; segOut = t2;
    ++level;
  }
  // phase 2
  GB(cudaMemcpyAsync(hostCounters, B.counters, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  GB(cudaStreamSynchronize(st));
  const uint32_t nSmall = hostCounters[2];
  const uint32_t lastLargeLevel = level ? level - 1 : 0;
  if (nSmall) k_finish_small<<<blocks(nSmall, 64), 64, 0, st>>>(nSmall, B.smLo, B.smHi, B.smNode, idxIn, B.blo, B.bhi, B.cen,
                                                                B.nodes, B.counters, maxDepth, splitMode);
  GB(cudaMemcpyAsync(hostCounters, B.counters, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  GB(cudaStreamSynchronize(st));
  GB(cudaGetLastError());
  const uint32_t nNodes = hostCounters[0];
  out.errorFlags = hostCounters[8];
  out.unsupported = hostCounters[3] != 0;
  out.depth = hostCounters[4] > lastLargeLevel ? hostCounters[4] : lastLargeLevel;
  if (out.unsupported || out.tooDeep || out.errorFlags) {    // caller falls back / reports
    cudaFreeAsync(arenaBase, st);
    return true;
  }

  // phase 3
  k_node_keys<<<blocks(nNodes), 256, 0, st>>>(nNodes, n, B.nodes, B.keysIn, B.idsIn, B.counters);
  GB(cub::DeviceRadixSort::SortPairs(B.sortTemp, sortBytes, B.keysIn, B.keysOut, B.idsIn, B.idsOut, (int)nNodes, 0, 64, st));
  GB(cudaMemcpyAsync(hostCounters, B.counters, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  GB(cudaStreamSynchronize(st));
  const uint32_t nInner = hostCounters[5], nMulti = hostCounters[6];
  out.nInner = nInner; out.nMulti = nMulti;
  // the five persistent arrays of the scene
  GB(cudaMalloc(&out.flat, (nInner ? nInner : 1) * 4 * sizeof(float4)));
  GB(cudaMalloc(&out.multi, (nMulti ? nMulti : 1) * sizeof(uint2)));
  GB(cudaMalloc(&out.prims, 3 * (size_t)n * sizeof(float4)));
  GB(cudaMalloc(&out.normals, 3 * (size_t)n * sizeof(float4)));
  GB(cudaMalloc(&out.order, (size_t)n * sizeof(uint32_t)));
  if (nInner) {
    k_pre_index<<<blocks(nInner), 256, 0, st>>>(nInner, B.idsOut, B.preIdx);
    k_emit_flat<<<blocks(nInner), 256, 0, st>>>(nInner, B.idsOut, B.nodes, B.preIdx, out.flat, out.multi, B.counters);
  }
  k_pack_prims<<<blocks(n), 256, 0, st>>>(in, idxIn, out.prims, out.normals);
  GB(cudaMemcpyAsync(out.order, idxIn, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  // root reference and box
  GpuNode root;
  GB(cudaMemcpyAsync(&root, B.nodes, sizeof(GpuNode), cudaMemcpyDeviceToHost, st));
  GB(cudaStreamSynchronize(st));
  GB(cudaGetLastError());
  for (int c = 0; c < 6; ++c) out.rootBox[c] = root.box[c];
  if (root.kind == 2) out.rootRef = 0;                                  // the root is first in pre-order
  else if (root.kind == 1) out.rootRef = 0x80000000u | root.lo;
  else {                                                                // a single multi-leaf
    const uint2 m = make_uint2(root.lo, root.hi - root.lo);
    GB(cudaMemcpy(out.multi, &m, sizeof(m), cudaMemcpyHostToDevice));
    out.rootRef = 0xC0000000u;
    out.nMulti = 1;
  }
  cudaFreeAsync(arenaBase, st);
  return true;
}

void freeGpuBuildOutput(GpuBuildOutput& out) {
  cudaFree(out.flat); cudaFree(out.multi); cudaFree(out.prims); cudaFree(out.normals); cudaFree(out.order);
  out.flat = nullptr; out.multi = nullptr; out.prims = nullptr; out.normals = nullptr; out.order = nullptr;
}

}  // namespace yb
