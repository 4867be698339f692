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

// Compressed copy of the wide nodes (64 B per node, child boxes on an 8-bit grid: conservative supersets) plus the exact
// boxes of the leaves, which the compressed walk tests when it visits a leaf (wide_bvh.cu, k_wide_compress).
struct CompressedWideOutput {
  float4* nodes = nullptr;       // 4 x float4 per wide node, same numbering as `wide`
  float4* leafBox = nullptr;     // 2 x float4 per primitive (DFS position): (lo.xyz, hi.x) (hi.yz, -, -)
  float4* multiBox = nullptr;    // the same per multi-leaf
  uint64_t bytes = 0;
  cudaError_t error = cudaSuccess;
  const char* where = "";
};
// Leaves `out.nodes` NULL (and returns true) when the tree cannot be compressed (non-finite box coordinates).
bool compressWideOnDevice(const float4* wide, uint32_t nWide, uint32_t nPrims, uint32_t nMulti, CompressedWideOutput& out);

}  // namespace yb
