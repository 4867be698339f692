// wide_bvh.cu -- collapses the reference's binary BVH (64-byte nodes, pre-order) into 4-wide nodes ON THE GPU.
//
// Why this keeps primitive IDs bit-exact.  In the reference a subtree is entered iff its box test passes with
// the ray's CURRENT tMax (wrapCollider, AABBs.hs:42-43; Culling.hs:24-25,33,38,52).  A parent's box is the
// `join` of its children's boxes (Culling.hs:39; min / max are exact), so every child box is contained in the
// parent box, and for a ray whose 1/u components are all finite the slab arithmetic (b - o) * (1/u) is monotone
// in b: the child's [tNear, tFar] interval lies inside the parent's.  Hence "child passes with tMax'" implies
// "parent passes with any tMax >= tMax'", and since tMax only shrinks during a walk:
//     a node is entered by the reference  <=>  its OWN box passes with the tMax current at that moment.
// The boxes of the ancestors are redundant; any hierarchy over the same leaf sequence (same left-first order,
// same leaf boxes) visits exactly the same leaves with exactly the same tMax values.  A wide node simply holds
// the boxes of up to four descendants of one binary node, in left-first order, and skips the boxes in between.
// Rays with a non-finite 1/u component (NaN slab products, SURVEY.md note N) are not monotone and keep using the
// binary tree.
//
//   k_wide_mark  : waves from the root; a marked binary node becomes a wide node, its expansion (below) names
//                  the binary nodes that become its wide children -> marked in the next wave
//   scan         : wide index = rank of the marked node in pre-order (keeps the left child adjacent)
//   k_wide_emit  : one thread per marked node writes the 128-byte wide node
//
// Expansion of binary node i: start with its two children; while fewer than four slots, replace the INNER slot
// with the largest surface area by its two children (in place, so the slots stay in left-first order).
#include "wide_bvh.hpp"

#include <cub/cub.cuh>

namespace yb {
namespace {

struct Slot {
  uint32_t ref;
  float lox, loy, hix, hiy, loz, hiz;
};

__device__ __forceinline__ void loadChildren(const float4* flat, uint32_t i, Slot& a, Slot& b) {
  const float4 n0 = flat[4 * (size_t)i + 0], n1 = flat[4 * (size_t)i + 1], n2 = flat[4 * (size_t)i + 2],
               n3 = flat[4 * (size_t)i + 3];
  a.ref = __float_as_uint(n3.x); a.lox = n0.x; a.loy = n0.y; a.hix = n0.z; a.hiy = n0.w; a.loz = n2.x; a.hiz = n2.y;
  b.ref = __float_as_uint(n3.y); b.lox = n1.x; b.loy = n1.y; b.hix = n1.z; b.hiy = n1.w; b.loz = n2.z; b.hiz = n2.w;
}

__device__ __forceinline__ bool isInner(uint32_t ref) { return (ref & kDevRefLeafBit) == 0; }   // null has the bit set

__device__ __forceinline__ float slotArea(const Slot& s) {
  const float dx = s.hix - s.lox, dy = s.hiy - s.loy, dz = s.hiz - s.loz;
  return dx * dy + dy * dz + dz * dx;
}

// The slots of the wide node rooted at binary node i, in left-first order.  Deterministic: the mark and the
// emit pass must agree.
__device__ int expandNode(const float4* flat, uint32_t i, Slot s[kWideWidth]) {
  int n = 2;
  loadChildren(flat, i, s[0], s[1]);
  while (n < kWideWidth) {
    int best = -1;
    float bestA = 0.0f;
    for (int k = 0; k < n; ++k) {
      if (!isInner(s[k].ref)) continue;
      const float a = slotArea(s[k]);
      if (best < 0 || a > bestA) { best = k; bestA = a; }
    }
    if (best < 0) break;
    for (int k = n; k > best + 1; --k) s[k] = s[k - 1];
    Slot l, r;
    loadChildren(flat, s[best].ref, l, r);
    s[best] = l;
    s[best + 1] = r;
    ++n;
  }
  return n;
}

__global__ void k_wide_mark(const float4* flat, uint32_t nInner, uint32_t* mark, uint32_t wave) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nInner || mark[i] != wave) return;
  Slot s[kWideWidth];
  const int n = expandNode(flat, i, s);
  for (int k = 0; k < n; ++k)
    if (isInner(s[k].ref)) mark[s[k].ref] = wave + 1u;
}

__global__ void k_wide_flags(const uint32_t* mark, uint32_t nInner, uint32_t* flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= nInner) flags[i] = (i < nInner && mark[i] != 0u) ? 1u : 0u;
}

__global__ void k_wide_emit(const float4* flat, uint32_t nInner, const uint32_t* mark, const uint32_t* wideIdx,
                            float4* wide) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nInner || mark[i] == 0u) return;
  Slot s[kWideWidth];
  const int n = expandNode(flat, i, s);
  float4* out = wide + kWideNodeVec * (size_t)wideIdx[i];
  uint32_t refs[kWideWidth];
  float z[2 * kWideWidth];
  for (int k = 0; k < kWideWidth; ++k) {
    if (k < n) {
      out[k] = make_float4(s[k].lox, s[k].loy, s[k].hix, s[k].hiy);
      z[2 * k] = s[k].loz; z[2 * k + 1] = s[k].hiz;
      refs[k] = isInner(s[k].ref) ? wideIdx[s[k].ref] : s[k].ref;
    } else {
      // empty slot: an inverted infinite box fails the slab test of every ray with finite 1/u
      out[k] = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
      z[2 * k] = INFINITY; z[2 * k + 1] = -INFINITY;
      refs[k] = kDevRefNull;
    }
  }
  out[4] = make_float4(z[0], z[1], z[2], z[3]);
  out[5] = make_float4(z[4], z[5], z[6], z[7]);
  out[6] = make_float4(__uint_as_float(refs[0]), __uint_as_float(refs[1]), __uint_as_float(refs[2]),
                       __uint_as_float(refs[3]));
  out[7] = make_float4(__uint_as_float((uint32_t)n), __uint_as_float(i), 0.0f, 0.0f);     // inspection only
}

// Deepest traversal stack the wide walk can need below a wide node: when child k is entered, at most the
// n - 1 - k later children are on the stack.  Runs bottom-up: children carry a later wave number.
__global__ void k_wide_need(const float4* wide, uint32_t nInner, const uint32_t* mark, const uint32_t* wideIdx,
                            uint32_t wave, uint32_t* need) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nInner || mark[i] != wave) return;
  const uint32_t w = wideIdx[i];
  const float4 rf = wide[kWideNodeVec * (size_t)w + 6];
  const uint32_t n = __float_as_uint(wide[kWideNodeVec * (size_t)w + 7].x);
  const uint32_t refs[4] = {__float_as_uint(rf.x), __float_as_uint(rf.y), __float_as_uint(rf.z), __float_as_uint(rf.w)};
  uint32_t m = 0;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t below = isInner(refs[k]) ? need[refs[k]] : 0u;
    m = max(m, (n - 1u - k) + below);
  }
  need[w] = m;
}

// COMPRESSED 4-wide nodes (64 B instead of 128 B: four 128-bit loads per visit instead of seven).
//
// Only the LEAF boxes have to be the reference's exact boxes: by the argument at the top of this file a node is entered
// by the reference iff its own box passes at that moment, so an inner box may be ANY superset of the exact one -- a
// superset only adds visits, and every visit ends in exact leaf tests.  A compressed node stores its child boxes on an
// 8-bit grid over the node's own box: per axis a grid step 2^e and an origin that is a multiple of the step;
// lo is rounded DOWN and hi UP to the grid, so the decoded box contains the exact one.  The decoded coordinate
// origin + q * 2^e is an exactly representable float by construction (|origin / 2^e| <= 2^23 - 512, q <= 255), the
// kernels evaluate it with one FMA as fma(2^23 + q, 2^e, origin - 2^23 * 2^e) -- exact operands, exact result, one
// rounding of a representable value -- and then run the SAME slab arithmetic on it as on an exact box: monotone in the
// box coordinates for a ray with finite 1/u, hence "exact child passes => decoded child passes" and the decoded entry
// distance is <= the exact one (it is the key the pop re-validates).  Leaf children are tested against their EXACT box
// (kept per DFS position in leafBox / per multi-leaf in multiBox) with the then-current tMax when the leaf is visited,
// which is precisely the reference's wrapCollider test (AABBs.hs:42-43).
//   node = 4 x float4: (origin' x y z, step exponents ex | ey << 8 | ez << 16) with origin' = origin - 2^23 * step,
//          (lo.x[4] lo.y[4] lo.z[4] hi.x[4]) and (hi.y[4] hi.z[4] - -) as bytes, the four child refs.
// A tree with a non-finite box coordinate, or an extent beyond 2^100, is not compressed (flag bad).
__global__ void k_wide_compress(const float4* wide, uint32_t nWide, float4* cw, float4* leafBox, float4* multiBox,
                                uint32_t* bad) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nWide) return;
  const float4* np = wide + kWideNodeVec * (size_t)w;
  const float4 z01 = np[4], z23 = np[5], rf = np[6];
  const uint32_t n = __float_as_uint(np[7].x);
  const uint32_t refs[4] = {__float_as_uint(rf.x), __float_as_uint(rf.y), __float_as_uint(rf.z), __float_as_uint(rf.w)};
  float lo[4][3], hi[4][3];
  for (int k = 0; k < 4; ++k) {
    const float4 b = np[k];
    lo[k][0] = b.x; lo[k][1] = b.y; hi[k][0] = b.z; hi[k][1] = b.w;
  }
  lo[0][2] = z01.x; hi[0][2] = z01.y; lo[1][2] = z01.z; hi[1][2] = z01.w;
  lo[2][2] = z23.x; hi[2][2] = z23.y; lo[3][2] = z23.z; hi[3][2] = z23.w;
  uint32_t qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0}, eb[3];
  float originp[3];
  for (int d = 0; d < 3; ++d) {
    double mn = INFINITY, mx = -INFINITY;
    for (uint32_t k = 0; k < n; ++k) { mn = fmin(mn, (double)lo[k][d]); mx = fmax(mx, (double)hi[k][d]); }
    if (!(isfinite(mn) && isfinite(mx)) || !(mn <= mx)) { *bad = 1u; return; }
    const double ext = mx - mn, mag = fmax(fabs(mn), fabs(mx));
    int e = -126;
    if (ext > 0.0) { int ee; frexp(ext / 253.0, &ee); e = max(e, ee); }            // 2^ee > ext / 253
    if (mag > 0.0) { int em; frexp(mag / 8388096.0, &em); e = max(e, em); }          // |origin / step| <= 2^23 - 512
    bool done = false;
    for (int attempt = 0; attempt < 4 && !done; ++attempt, ++e) {
      if (e > 100) { *bad = 1u; return; }
      const double step = ldexp(1.0, e);
      const double origin = floor(mn / step) * step;                 // a multiple of the step, <= mn
      uint32_t wl = 0, wh = 0;
      bool fits = true;
      for (uint32_t k = 0; k < 4 && fits; ++k) {
        uint32_t a = 255u, b = 0u;                                   // empty slot: inverted box
        if (k < n) {
          // floor / ceil of (x - origin) / step; x - origin is NOT exact in binary64 when the magnitudes are far apart
          // (x = 1e-16, origin = -8), so the candidates are corrected against the decoded values: origin + q * step IS
          // exact (it is a binary32 number), and so are the comparisons with the exact coordinates
          double qa = floor(((double)lo[k][d] - origin) / step), qb = ceil(((double)hi[k][d] - origin) / step);
          while (origin + qa * step > (double)lo[k][d]) qa -= 1.0;
          while (origin + qb * step < (double)hi[k][d]) qb += 1.0;
          if (qa < 0.0 || qb > 255.0 || qa > qb) { fits = false; break; }
          a = (uint32_t)qa; b = (uint32_t)qb;
        }
        wl |= a << (8u * k);
        wh |= b << (8u * k);
      }
      if (!fits) continue;                                           // one grid step too fine: double it
      qlo[d] = wl; qhi[d] = wh;
      eb[d] = (uint32_t)(e + 127);
      originp[d] = (float)(origin - 8388608.0 * step);
      done = true;
    }
    if (!done) { *bad = 1u; return; }
  }
  float4* out = cw + 4 * (size_t)w;
  out[0] = make_float4(originp[0], originp[1], originp[2], __uint_as_float(eb[0] | (eb[1] << 8) | (eb[2] << 16)));
  out[1] = make_float4(__uint_as_float(qlo[0]), __uint_as_float(qlo[1]), __uint_as_float(qlo[2]), __uint_as_float(qhi[0]));
  out[2] = make_float4(__uint_as_float(qhi[1]), __uint_as_float(qhi[2]), 0.0f, 0.0f);
  out[3] = rf;
  for (uint32_t k = 0; k < n; ++k) {                                // exact boxes of the leaf children
    const uint32_t r = refs[k];
    if (r == kDevRefNull || !(r & kDevRefLeafBit)) continue;
    float4* dst = ((r & kDevRefMultiBits) == kDevRefMultiBits ? multiBox : leafBox) + 2 * (size_t)(r & 0x3FFFFFFFu);
    dst[0] = make_float4(lo[k][0], lo[k][1], lo[k][2], hi[k][0]);
    dst[1] = make_float4(hi[k][1], hi[k][2], 0.0f, 0.0f);
  }
}

inline unsigned blocks(size_t n, unsigned per = 256) { return (unsigned)((n + per - 1) / per); }

}  // namespace

#define WB(call)                                                                            \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      out.error = e__; out.where = #call;                                                   \
      if (arena) cudaFreeAsync(arena, st);                                                  \
      cudaFree(out.wide);                                                                   \
      out.wide = nullptr;                                                                   \
      return false;                                                                         \
    }                                                                                       \
  } while (0)

bool compressWideOnDevice(const float4* wide, uint32_t nWide, uint32_t nPrims, uint32_t nMulti, CompressedWideOutput& out) {
  out = CompressedWideOutput();
  if (!wide || nWide == 0) return true;
  uint32_t* bad = nullptr;
  auto failWith = [&](cudaError_t e, const char* where) {
    out.error = e; out.where = where;
    cudaFree(out.nodes); cudaFree(out.leafBox); cudaFree(out.multiBox); cudaFree(bad);
    out.nodes = out.leafBox = out.multiBox = nullptr;
    return false;
  };
  cudaError_t e;
  if ((e = cudaMalloc(&out.nodes, (size_t)nWide * 4 * sizeof(float4))) != cudaSuccess) return failWith(e, "cudaMalloc(cwide)");
  if ((e = cudaMalloc(&out.leafBox, (size_t)(nPrims ? nPrims : 1) * 2 * sizeof(float4))) != cudaSuccess) return failWith(e, "cudaMalloc(leafBox)");
  if ((e = cudaMalloc(&out.multiBox, (size_t)(nMulti ? nMulti : 1) * 2 * sizeof(float4))) != cudaSuccess) return failWith(e, "cudaMalloc(multiBox)");
  if ((e = cudaMalloc(&bad, sizeof(uint32_t))) != cudaSuccess) return failWith(e, "cudaMalloc(flag)");
  if ((e = cudaMemset(bad, 0, sizeof(uint32_t))) != cudaSuccess) return failWith(e, "cudaMemset(flag)");
  k_wide_compress<<<blocks(nWide), 256>>>(wide, nWide, out.nodes, out.leafBox, out.multiBox, bad);
  uint32_t flag = 0;
  if ((e = cudaMemcpy(&flag, bad, sizeof(flag), cudaMemcpyDeviceToHost)) != cudaSuccess) return failWith(e, "k_wide_compress");
  cudaFree(bad);
  bad = nullptr;
  if (flag) {                    // not compressible (non-finite or huge boxes): the scene keeps the exact 128-byte nodes
    cudaFree(out.nodes); cudaFree(out.leafBox); cudaFree(out.multiBox);
    out.nodes = out.leafBox = out.multiBox = nullptr;
    return true;
  }
  out.bytes = (uint64_t)nWide * 64 + (uint64_t)nPrims * 32 + (uint64_t)nMulti * 32;
  return true;
}

bool buildWideOnDevice(const float4* flat, uint32_t nInner, uint32_t binaryDepth, WideBuildOutput& out) {
  out = WideBuildOutput();
  if (nInner == 0 || !flat) return true;              // the root is a leaf (or the scene is empty): nothing to collapse
  cudaStream_t st = nullptr;
  char* arena = nullptr;
  // temporaries come from one stream-ordered allocation (the pool keeps its memory between builds)
  size_t tempBytes = 0;
  WB(cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(nInner + 1), st));
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t words = align(((size_t)nInner + 1) * sizeof(uint32_t));
  const size_t total = 4 * words + align(tempBytes ? tempBytes : 1);
  WB(cudaMallocAsync((void**)&arena, total, st));
  uint32_t* mark = reinterpret_cast<uint32_t*>(arena);
  uint32_t* flags = reinterpret_cast<uint32_t*>(arena + words);
  uint32_t* scan = reinterpret_cast<uint32_t*>(arena + 2 * words);
  uint32_t* need = reinterpret_cast<uint32_t*>(arena + 3 * words);       // nWide <= nInner entries
  void* temp = arena + 4 * words;
  WB(cudaMemsetAsync(mark, 0, (size_t)nInner * sizeof(uint32_t), st));
  const uint32_t one = 1u;
  WB(cudaMemcpyAsync(mark, &one, sizeof(one), cudaMemcpyHostToDevice, st));     // the root (pre-order index 0)
  // every wave descends at least one binary level, so binaryDepth + 2 waves reach every node
  const uint32_t nWaves = binaryDepth + 2u;
  for (uint32_t wave = 1; wave <= nWaves; ++wave)
    k_wide_mark<<<blocks(nInner), 256, 0, st>>>(flat, nInner, mark, wave);
  k_wide_flags<<<blocks((size_t)nInner + 1), 256, 0, st>>>(mark, nInner, flags);
  WB(cub::DeviceScan::ExclusiveSum(temp, tempBytes, flags, scan, (int)(nInner + 1), st));
  uint32_t nWide = 0;
  WB(cudaMemcpyAsync(&nWide, scan + nInner, sizeof(nWide), cudaMemcpyDeviceToHost, st));
  WB(cudaStreamSynchronize(st));
  WB(cudaGetLastError());
  WB(cudaMalloc(&out.wide, (size_t)(nWide ? nWide : 1) * kWideNodeVec * sizeof(float4)));
  k_wide_emit<<<blocks(nInner), 256, 0, st>>>(flat, nInner, mark, scan, out.wide);
  for (uint32_t wave = nWaves + 1u; wave >= 1u; --wave)
    k_wide_need<<<blocks(nInner), 256, 0, st>>>(out.wide, nInner, mark, scan, wave, need);
  uint32_t rootNeed = 0;
  WB(cudaMemcpyAsync(&rootNeed, need, sizeof(rootNeed), cudaMemcpyDeviceToHost, st));
  WB(cudaStreamSynchronize(st));
  WB(cudaGetLastError());
  out.nWide = nWide;
  out.stackNeed = rootNeed;
  cudaFreeAsync(arena, st);
  return true;
}

}  // namespace yb
