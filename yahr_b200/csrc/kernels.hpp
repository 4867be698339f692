// kernels.hpp -- host-callable launchers of the CUDA kernels.
#pragma once
#include "device_types.cuh"

namespace yb {

// One CTA per tile, one thread per pixel, whole per-pixel path inline.
cudaError_t launchRenderMega(const RenderParams& P, cudaStream_t stream, uint32_t* launches);

// float RGB -> 8-bit RGB with JuicyPixels' quantisation, for elements [first, first + count).
cudaError_t launchQuantizeRgb8(const float* rgb, unsigned char* out, size_t first, size_t count, cudaStream_t stream,
                               uint32_t* launches);

// Multi-GPU end-of-frame fence: one word per pushing rank in rank 0's memory.
cudaError_t launchFlagSignal(uint32_t* flag, uint32_t value, cudaStream_t stream);
cudaError_t launchFlagsWait(uint32_t* flags, int count, uint32_t value, cudaStream_t stream);

}  // namespace yb

namespace yb {
// Persistent-thread wavefront kernel set (primary trace, shade, shadow trace, resolve); depth 1 only.
// phaseEvents: NULL or 4 events recorded before primary / shade / shadow and after shadow (first sample).
// Fills table[item] = u | v << 16 for every item of the tile set described by W (tiles, tileStart, nItems).
cudaError_t launchPixelTable(WavefrontParams W, uint32_t* table, cudaStream_t stream);
cudaError_t launchWavefront(WavefrontParams W, int numSMs, cudaStream_t stream, uint32_t* launches,
                            cudaEvent_t* phaseEvents);
// The counting build of the same kernels (wavefront_count.cu): depth 1 only; fills W.workStats.
namespace counted {
cudaError_t launchWavefrontCounted(WavefrontParams W, int numSMs, cudaStream_t stream, uint32_t* launches,
                                   cudaEvent_t* phaseEvents);
}
}  // namespace yb
