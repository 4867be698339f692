// wide_bvh.hpp -- 4-wide collapse of the reference's binary BVH (see wide_bvh.cu for the exactness argument).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "device_types.cuh"

namespace yb {

struct WideBuildOutput {
  float4* wide = nullptr;        // kWideNodeVec x float4 per wide node, numbered in pre-order of their binary roots
  uint32_t nWide = 0;
  uint32_t stackNeed = 0;        // deepest traversal stack the wide walk can need (entries)
  cudaError_t error = cudaSuccess;
  const char* where = "";
};

// flat: the device-resident binary nodes (4 x float4 each, pre-order, root = 0); binaryDepth: depth of that tree.
// Returns false on a CUDA error.  nInner == 0 leaves out.wide NULL (the root reference is a leaf or null).
bool buildWideOnDevice(const float4* flat, uint32_t nInner, uint32_t binaryDepth, WideBuildOutput& out);

}  // namespace yb
