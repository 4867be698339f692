// kernels.hpp -- host-callable launchers of the CUDA kernels.
#pragma once
#include "device_types.cuh"

namespace yb {

// One CTA per tile, one thread per pixel, whole per-pixel path inline.
cudaError_t launchRenderMega(const RenderParams& P, cudaStream_t stream, uint32_t* launches);

}  // namespace yb
