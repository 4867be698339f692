// wavefront.cu -- persistent-thread wavefront kernel set (recursion depth 1 = direct lighting, the
// mode of every BASELINE config).  Compile with --fmad=false.
//
//   k_wf_primary : persistent warps pull work items (pixels in tile-major order) from a global counter,
//                  generate the camera ray (A3) and run the closest-hit walk (A4-A8).  Lanes whose ray
//                  has finished are refilled with new items (ballot + popc compaction, one atomic per
//                  warp) so the traversal loop keeps running with a full warp.  Output: (t, DFS position)
//                  per item.
//   k_wf_shade   : one thread per item: DifferentialGeometry, material, BSDF per light (A9-A13);
//                  writes the pixel's base value and appends one shadow-probe record per light with a
//                  non-zero BSDF (Integrators.hs:59) to a queue (warp-aggregated append).
//   k_wf_shadow  : persistent warps with the same refill scheme run any-hit walks over the compacted
//                  queue; an unoccluded probe adds its contribution to the pixel (single light) or sets
//                  its visibility flag (several lights; k_wf_resolve sums them in light order).
//   k_wf_accum   : spp > 1 only: acc += sample, and the final divide.
//
// The walk is the reference's (left child first, test on entry with the current tMax, later hit
// replaces), organised as while-while: all lanes descend inner nodes until each holds a leaf, then the
// leaf code runs once for the whole warp.
#include "kernels.hpp"
#include "render_device.cuh"

namespace yb {

using namespace dev;

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kRefillThreshold = 20;   // leave the traversal loop when fewer lanes than this are busy

struct Trav {
  uint32_t cur;
  int sp;
  float tMax;
  uint32_t best;
};

// Pops the next deferred subtree whose box test still passes with the current tMax.
__device__ __forceinline__ bool popNext(Trav& s, const uint2* stack) {
  while (s.sp > 0) {
    const uint2 e = stack[--s.sp];
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

// One while-while round.  Returns true when the walk is finished.
template <bool ANY_HIT, bool ORDERED>
__device__ __forceinline__ bool travRound(const DeviceScene& sc, const Ray& r, Trav& s, uint2* stack) {
  // ---- descend inner nodes until this lane holds a leaf -------------------------------------
  while (!(s.cur & kDevRefLeafBit)) {
    const float4* np = sc.nodes + 4 * (size_t)s.cur;
    const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
    float keyL, keyR;
    const uint32_t refL = __float_as_uint(n3.x), refR = __float_as_uint(n3.y);
    const bool passL = boxTest(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, r, s.tMax, keyL) && refL != kDevRefNull;
    const bool passR = boxTest(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, r, s.tMax, keyR) && refR != kDevRefNull;
    if (passL && passR) {
      if (ORDERED && keyR < keyL) {
        stack[s.sp++] = make_uint2(refL, __float_as_uint(keyL));
        s.cur = refR;
      } else {
        stack[s.sp++] = make_uint2(refR, __float_as_uint(keyR));
        s.cur = refL;
      }
    } else if (passL) {
      s.cur = refL;
    } else if (passR) {
      s.cur = refR;
    } else if (!popNext(s, stack)) {
      return true;
    }
  }
  // ---- leaf ----------------------------------------------------------------------------------
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
    if (hitPrimitive(sc, idx, r, s.tMax, t)) {
      if (!ORDERED || t < s.tMax || s.best == kNoHit || idx > s.best) {
        s.best = idx; s.tMax = t;
        if (ANY_HIT) return true;
      }
    }
  }
  return !popNext(s, stack);
}

// item index (tile-major) -> tile.  Tiles are almost uniform in size, so a proportional guess is
// off by a step or two at most.
__device__ __forceinline__ uint32_t tileOfItem(const WavefrontParams& W, uint32_t item) {
  uint32_t t = (uint32_t)(((unsigned long long)item * W.base.nTiles) / W.nItems);
  while (__ldg(&W.tileStart[t]) > item) --t;
  while (__ldg(&W.tileStart[t + 1]) <= item) ++t;
  return t;
}

__device__ __forceinline__ void itemPixel(const WavefrontParams& W, uint32_t item, int& u, int& v) {
  const uint32_t t = tileOfItem(W, item);
  const int4 win = __ldg(&W.base.tiles[t]);
  const int i = (int)(item - __ldg(&W.tileStart[t]));
  const int th = win.w - win.y;
  u = win.x + i / th;                                 // u-major inside the tile (Sampling.hs:6)
  v = win.y + i % th;
}

__device__ __forceinline__ Ray itemRay(const WavefrontParams& W, int u, int v, uint32_t sample) {
  const uint32_t pixel = (uint32_t)(W.base.width * v + u);
  const float fu = (float)u + sampleOffset(W.base.seed, pixel, sample, 0);
  const float fv = (float)v + sampleOffset(W.base.seed, pixel, sample, 1);
  return cameraRay(W.base, fu, fv);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
template <bool ORDERED>
__global__ void __launch_bounds__(128) k_wf_primary(const __grid_constant__ WavefrontParams W) {
  uint2 stack[64];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned ltMask = (1u << lane) - 1u;
  bool busy = false, exhausted = false;      // exhausted is warp-uniform
  uint32_t item = 0;
  Ray r;
  Trav s;
  s.cur = 0; s.sp = 0; s.tMax = 0; s.best = kNoHit;
  for (;;) {
    // ---- refill idle lanes with new work items ----------------------------------------------
    const unsigned idle = __ballot_sync(kFull, !busy);
    if (idle && !exhausted) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&W.work[0], (uint32_t)__popc(idle));
      base = __shfl_sync(kFull, base, 0);
      exhausted = base + (uint32_t)__popc(idle) >= W.nItems;
      if (!busy) {
        const uint32_t mine = base + __popc(idle & ltMask);
        if (mine < W.nItems) {
          item = mine;
          int u, v;
          itemPixel(W, item, u, v);
          r = itemRay(W, u, v, W.sample);
          busy = travBegin(W.base.sc, r, 1e6f, s);
          if (!busy) { W.hitT[item] = 0.0f; W.hitIdx[item] = kNoHit; }
        }
      }
    }
    if (!__any_sync(kFull, busy)) {
      if (exhausted) break;                   // queue drained and nothing in flight
      continue;
    }
    // ---- traverse until too few lanes are busy ----------------------------------------------
    const int threshold = exhausted ? 1 : kRefillThreshold;
    do {
      if (busy) {
        if (travRound<false, ORDERED>(W.base.sc, r, s, stack)) {
          W.hitT[item] = s.tMax;
          W.hitIdx[item] = s.best;
          busy = false;
        }
      }
    } while (__popc(__ballot_sync(kFull, busy)) >= threshold);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_shade(const __grid_constant__ WavefrontParams W) {
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  const bool valid = item < W.nItems;
  const DeviceScene& sc = W.base.sc;
  V3 x = mk(0, 0, 0);
  Surface surf;
  Frame fr;
  MaterialD mat;
  uint32_t pixel = 0;
  bool hit = false;
  uint32_t nanBits = 0;
  V3 wo = mk(0, 0, 0);
  if (valid) {
    int u, v;
    itemPixel(W, item, u, v);
    pixel = (uint32_t)(W.base.width * v + u);
    const uint32_t idx = W.hitIdx[item];
    float* out = W.sampleOut + 3 * (size_t)pixel;
    if (idx == kNoHit) {
      out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f;
      if (W.base.primid && W.sample == 0) W.base.primid[pixel] = kNoHit;
    } else {
      hit = true;
      const Ray r = itemRay(W, u, v, W.sample);
      surf = surfaceAt(sc, idx, r, W.hitT[item]);
      if (W.base.primid && W.sample == 0) W.base.primid[pixel] = surf.primId;
      mat = loadMaterial(sc, surf.material);
      fr = makeFrame(surf);
      wo = vneg(r.d);
      x = surf.x;
      // ((n . r) @* f r) * rs with rs = vcast 0 = 0 (Integrators.hs:26,37,41-43): +-0, or NaN when the
      // weight is not finite.  Stored as the pixel's base value; the direct term is added to it.
      const V3 refl = vsub(r.d, vscale(2.0f * dot(r.d, surf.n), surf.n));
      const V3 w = vscale(dot(surf.n, refl), bsdfAt(mat, fr, refl, wo));
      const V3 base = vadd(vmul(w, mk(0.0f, 0.0f, 0.0f)), mk(0.0f, 0.0f, 0.0f));
      out[0] = base.x; out[1] = base.y; out[2] = base.z;
      nanBits = (base.x != base.x ? 1u : 0u) | (base.y != base.y ? 2u : 0u) | (base.z != base.z ? 4u : 0u);
    }
  }
  // ---- emit shadow probes: one per light with lensq k > 0 ------------------------------------
  // single light: compacted queue (warp-aggregated append).  several lights: dense slots
  // entry = item * nLights + light so that k_wf_resolve can sum a pixel's lights in order.
  uint32_t nEmit = 0;
  for (uint32_t li = 0; li < sc.nLights; ++li) {
    bool emit = false;
    V3 p0 = mk(0, 0, 0), dl = mk(0, 0, 0), contrib = mk(0, 0, 0);
    if (hit) {
      const V3 lightPos = xyz(__ldg(&sc.lights[2 * li + 0]));
      const V3 spectrum = xyz(__ldg(&sc.lights[2 * li + 1]));
      const V3 pointToLight = vsub(lightPos, x);
      const V3 lightDir = vnorm(pointToLight);
      const V3 k = bsdfAt(mat, fr, lightDir, wo);
      if (lensq(k) > 0.0f) {
        emit = true;
        p0 = vadd(x, vscale(0.001f, lightDir));
        dl = vsub(lightPos, p0);
        const V3 intensity = vscale(rcp(lensq(pointToLight)), spectrum);
        contrib = vmul(vscale(fabsf(dot(lightDir, surf.n)), k), intensity);
      }
    }
    uint32_t e = 0;
    if (W.dense) {
      e = item * sc.nLights + li;
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
      W.q1[e] = make_float4(d.x, d.y, d.z, __uint_as_float(pixel));
      W.q2[e] = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(nanBits));
      ++nEmit;
    }
  }
  // shadow-ray count for the stats
  const uint32_t warpEmit = __reduce_add_sync(kFull, nEmit);
  if (lane == 0 && warpEmit) atomicAdd(&W.work[3], warpEmit);
}

// ---------------------------------------------------------------------------------------------
// Result of one shadow probe.  Single light: the pixel holds its base value (+-0 or NaN, written by
// k_wf_shade); an unoccluded probe overwrites it with base + (0 + contribution), which is the
// contribution itself unless the base is NaN (flag bits in q2.w) -- no read-modify-write, so the
// frame may live in a peer GPU.  Several lights: only the visibility flag is recorded.
__device__ __forceinline__ void shadowResult(const WavefrontParams& W, uint32_t entry, bool unoccluded) {
  if (W.visibility) { W.visibility[entry] = unoccluded ? 1 : 0; return; }
  if (!unoccluded) return;
  const float4 b = W.q1[entry], c = W.q2[entry];
  const uint32_t nanBits = __float_as_uint(c.w);
  const float qnan = __uint_as_float(0x7FFFFFFFu);
  float* out = W.sampleOut + 3 * (size_t)__float_as_uint(b.w);
  out[0] = (nanBits & 1u) ? qnan : 0.0f + c.x;
  out[1] = (nanBits & 2u) ? qnan : 0.0f + c.y;
  out[2] = (nanBits & 4u) ? qnan : 0.0f + c.z;
}

template <bool ORDERED>
__global__ void __launch_bounds__(128) k_wf_shadow(const __grid_constant__ WavefrontParams W) {
  uint2 stack[64];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned ltMask = (1u << lane) - 1u;
  const uint32_t nEntries = W.dense ? W.nItems * W.base.sc.nLights : W.work[2];
  bool busy = false, exhausted = false;
  uint32_t entry = 0;
  Ray r;
  Trav s;
  s.cur = 0; s.sp = 0; s.tMax = 0; s.best = kNoHit;
  for (;;) {
    const unsigned idle = __ballot_sync(kFull, !busy);
    if (idle && !exhausted) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&W.work[1], (uint32_t)__popc(idle));
      base = __shfl_sync(kFull, base, 0);
      exhausted = base + (uint32_t)__popc(idle) >= nEntries;
      if (!busy) {
        const uint32_t mine = base + __popc(idle & ltMask);
        if (mine < nEntries) {
          entry = mine;
          const float4 a = W.q0[entry], b = W.q1[entry];
          if (a.w < 0.0f) {                      // empty dense slot
            W.visibility[entry] = 0;
          } else {
            r = makeRay(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z));
            busy = travBegin(W.base.sc, r, a.w, s);
            if (!busy) shadowResult(W, entry, true);   // the root box already rejects the probe
          }
        }
      }
    }
    if (!__any_sync(kFull, busy)) {
      if (exhausted) break;
      continue;
    }
    const int threshold = exhausted ? 1 : kRefillThreshold;
    do {
      if (busy) {
        if (travRound<true, ORDERED>(W.base.sc, r, s, stack)) {
          busy = false;
          shadowResult(W, entry, s.best == kNoHit);
        }
      }
    } while (__popc(__ballot_sync(kFull, busy)) >= threshold);
  }
}

// Several lights: sum the visible contributions of a pixel in light order (sum = foldl (+) 0).
__global__ void __launch_bounds__(256) k_wf_resolve(const __grid_constant__ WavefrontParams W) {
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= W.nItems) return;
  const uint32_t nL = W.base.sc.nLights;
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

// spp > 1: acc = acc + sample (sample order), then pixel = acc / spp on the last sample.
__global__ void __launch_bounds__(256) k_wf_accum(const __grid_constant__ WavefrontParams W) {
  const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= W.nItems) return;
  int u, v;
  itemPixel(W, item, u, v);
  const size_t p = 3 * (size_t)(W.base.width * v + u);
  const float n = (float)W.base.spp;
  for (int c = 0; c < 3; ++c) {
    const float a = (W.sample == 0 ? 0.0f : W.accum[p + c]) + W.sampleOut[p + c];
    if ((int)W.sample == W.base.spp - 1) W.base.rgb[p + c] = __fdiv_rn(a, n);
    else W.accum[p + c] = a;
  }
}

__global__ void k_wf_count(const __grid_constant__ WavefrontParams W) {
  // ray counters for the stats: primary = items, shadow = queue length
  atomicAdd(&W.base.counters[0], (unsigned long long)W.nItems);
  atomicAdd(&W.base.counters[1], (unsigned long long)W.work[3]);
}

cudaError_t launchWavefront(WavefrontParams W, int numSMs, cudaStream_t stream, uint32_t* launches,
                            cudaEvent_t* phaseEvents) {
  if (W.nItems == 0) return cudaSuccess;
  const bool ordered = W.base.traversal == 1;
  const int persistentBlocks = numSMs * 8;             // 128-thread CTAs, up to 8 resident per SM
  const uint32_t itemBlocks = (W.nItems + 255u) / 256u;
  for (int s = 0; s < W.base.spp; ++s) {
    W.sample = (uint32_t)s;
    W.sampleOut = W.base.spp == 1 ? W.base.rgb : W.sampleBuf;
    cudaError_t e = cudaMemsetAsync(W.work, 0, 4 * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const bool timed = phaseEvents && s == 0;      // phase times of the first sample pass
    if (timed) cudaEventRecord(phaseEvents[0], stream);
    if (ordered) k_wf_primary<true><<<persistentBlocks, 128, 0, stream>>>(W);
    else k_wf_primary<false><<<persistentBlocks, 128, 0, stream>>>(W);
    if (timed) cudaEventRecord(phaseEvents[1], stream);
    k_wf_shade<<<itemBlocks, 256, 0, stream>>>(W);
    if (timed) cudaEventRecord(phaseEvents[2], stream);
    if (ordered) k_wf_shadow<true><<<persistentBlocks, 128, 0, stream>>>(W);
    else k_wf_shadow<false><<<persistentBlocks, 128, 0, stream>>>(W);
    if (timed) cudaEventRecord(phaseEvents[3], stream);
    if (launches) *launches += 3;
    if (W.dense) { k_wf_resolve<<<itemBlocks, 256, 0, stream>>>(W); if (launches) *launches += 1; }
    if (W.base.spp > 1) { k_wf_accum<<<itemBlocks, 256, 0, stream>>>(W); if (launches) *launches += 1; }
    k_wf_count<<<1, 1, 0, stream>>>(W);
    if (launches) *launches += 1;
  }
  return cudaGetLastError();
}

}  // namespace yb
