// kernels.cu -- sm_100a kernels of libyahr_b200.  Compile with --fmad=false (bit-exact primitive IDs).
#include "kernels.hpp"
#include "render_device.cuh"

namespace yb {

using namespace dev;

// ---------------------------------------------------------------------------------------------
// Kernel "mega": one CTA per reference tile (the reference's unit of parallelism: one spark /
// Par task per squareBatches window, main.hs:82,93), one thread per pixel of the tile, the whole
// per-pixel path `li (cast u v)` (main.hs:73) inline: camera ray -> closest hit -> shade ->
// shadow probes -> (reflection levels) -> pixel store at w*v+u (main.hs:100).
// ---------------------------------------------------------------------------------------------
template <bool ORDERED>
__global__ void __launch_bounds__(256) k_render_mega(const __grid_constant__ RenderParams P) {
  const int4 win = P.tiles[blockIdx.x];
  const int th = win.w - win.y;
  const int npx = (win.z - win.x) * th;
  Counters cnt;
  cnt.primary = cnt.shadow = cnt.secondary = 0;
  for (int i = threadIdx.x; i < npx; i += blockDim.x) {
    const int u = win.x + i / th, v = win.y + i % th;   // [(u, v) | u <- [x0..x1-1], v <- [y0..y1-1]]
    const uint32_t pixel = (uint32_t)(P.width * v + u);
    V3 acc = mk(0.0f, 0.0f, 0.0f), L = acc;
    uint32_t prim0 = kNoHit;
    for (int s = 0; s < P.spp; ++s) {
      const float fu = (float)u + sampleOffset(P.seed, pixel, (uint32_t)s, 0);
      const float fv = (float)v + sampleOffset(P.seed, pixel, (uint32_t)s, 1);
      const Ray ray = cameraRay(P, fu, fv);
      cnt.primary++;
      uint32_t prim;
      ShadeCtx ctx;
      ctx.seed = P.seed; ctx.pixel = pixel; ctx.sample = (uint32_t)s; ctx.level = 0;
      L = radiance<ORDERED>(P, ray, cnt, prim, ctx);
      if (s == 0) prim0 = prim;
      acc = vadd(acc, L);
    }
    if (P.spp != 1) {
      const float n = (float)P.spp;
      L = mk(__fdiv_rn(acc.x, n), __fdiv_rn(acc.y, n), __fdiv_rn(acc.z, n));
    }
    float* out = P.rgb + 3 * (size_t)pixel;
    out[0] = L.x; out[1] = L.y; out[2] = L.z;
    if (P.primid) P.primid[pixel] = prim0;
  }
  // ray counters: one atomic per warp per counter
  const unsigned full = 0xFFFFFFFFu;
  const uint32_t p = __reduce_add_sync(full, cnt.primary);
  const uint32_t sh = __reduce_add_sync(full, cnt.shadow);
  const uint32_t se = __reduce_add_sync(full, cnt.secondary);
  if ((threadIdx.x & 31) == 0) {
    if (p) atomicAdd(&P.counters[0], (unsigned long long)p);
    if (sh) atomicAdd(&P.counters[1], (unsigned long long)sh);
    if (se) atomicAdd(&P.counters[2], (unsigned long long)se);
  }
}

// JuicyPixels' ImageRGBF -> 8-bit conversion of savePngImage (main.hs:142): truncate (255 * max 0 (min 1 x)),
// no gamma; GHC's min/max selects turn NaN into 0.
__global__ void __launch_bounds__(256) k_quantize_rgb8(const float* __restrict__ rgb, unsigned char* __restrict__ out,
                                                        size_t first, size_t count) {
  const size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= first + count) return;
  const float x = rgb[i];
  const float m = (1.0f <= x) ? 1.0f : x;
  const float c = (0.0f <= m) ? m : 0.0f;
  out[i] = (unsigned char)(int)(255.0f * c);
}

cudaError_t launchQuantizeRgb8(const float* rgb, unsigned char* out, size_t first, size_t count, cudaStream_t stream,
                               uint32_t* launches) {
  if (count == 0) return cudaSuccess;
  k_quantize_rgb8<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(rgb, out, first, count);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// Multi-GPU end-of-frame fence (yahr_b200_flag_signal / yahr_b200_flags_wait).  The signalling kernel runs after the
// rank's stores / copies into the gather frame in stream order; the fence makes them visible system-wide before the word.
__global__ void k_flag_signal(uint32_t* flag, uint32_t value) {
  __threadfence_system();
  *(volatile uint32_t*)flag = value;
}

// One thread per word; sequence numbers only grow (wrap-safe comparison).  A peer that never signals must not hang the
// device: after ~4 s the wait gives up (the frame is then incomplete and the word is left behind, which the next wait
// and the caller's frame check will show).
__global__ void k_flags_wait(uint32_t* flags, int count, uint32_t value) {
  const int i = (int)threadIdx.x;
  if (i >= count) return;
  const long long t0 = clock64();
  while ((int)(*(volatile uint32_t*)(flags + i) - value) < 0) {
    __nanosleep(100);
    if (clock64() - t0 > 8000000000ll) break;
  }
  __threadfence_system();
}

cudaError_t launchFlagSignal(uint32_t* flag, uint32_t value, cudaStream_t stream) {
  k_flag_signal<<<1, 1, 0, stream>>>(flag, value);
  return cudaGetLastError();
}

cudaError_t launchFlagsWait(uint32_t* flags, int count, uint32_t value, cudaStream_t stream) {
  k_flags_wait<<<1, count <= 32 ? 32 : ((count + 31) / 32) * 32, 0, stream>>>(flags, count, value);
  return cudaGetLastError();
}

cudaError_t launchRenderMega(const RenderParams& P, cudaStream_t stream, uint32_t* launches) {
  if (P.nTiles == 0) return cudaSuccess;
  if (P.traversal == 1) k_render_mega<true><<<P.nTiles, 256, 0, stream>>>(P);
  else k_render_mega<false><<<P.nTiles, 256, 0, stream>>>(P);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace yb
