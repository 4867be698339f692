// bvh_build_gpu.hpp -- device-side construction of the reference's BVH (see bvh_build_gpu.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace yb {

// The caller's primitive arrays, already copied to the device, in the layout of yahr_scene_desc.
struct GpuBuildInput {
  uint32_t nPrims, nTriangles, nSpheres, nMaterials;
  const float *triP0, *triP1, *triP2, *triN0, *triN1, *triN2;
  const uint32_t* triMaterial;
  const float *sphCenter, *sphRadius;
  const uint32_t* sphMaterial;
  const uint32_t* primOrder;     // may be null: spheres then triangles
};

struct GpuNode {                 // reference-tree node during construction
  uint32_t kind;                 // 1 single leaf, 2 inner, 3 multi-leaf
  uint32_t lo, hi;               // range of DFS positions
  int32_t left, right;
  uint32_t level;
  float box[6];
};

struct GpuBuildOutput {
  // device arrays in the traversal layout (device_types.cuh); ownership passes to the caller
  float4* flat = nullptr;        // 4 x float4 per inner node, pre-order numbering
  uint2* multi = nullptr;
  float4* prims = nullptr;       // 3 x float4 per primitive, DFS order
  float4* normals = nullptr;
  uint32_t* order = nullptr;     // primitive ID per DFS position
  uint32_t nInner = 0, nMulti = 0, depth = 0;
  uint32_t rootRef = 0xFFFFFFFFu;
  float rootBox[6] = {0, 0, 0, 0, 0, 0};
  uint32_t errorFlags = 0;       // 1 non-finite geometry, 2 material index out of range, 4 bad prim_order entry
  bool unsupported = false;      // a case left to the host builder (empty left partition)
  bool tooDeep = false;
  cudaError_t error = cudaSuccess;
  const char* where = "";
};

bool gpuBuildSupported(int splitMode);
// Returns false on a CUDA error (out.error / out.where).  On success check errorFlags / unsupported / tooDeep.
bool buildBvhOnDevice(const GpuBuildInput& in, int maxDepth, int splitMode, GpuBuildOutput& out);
void freeGpuBuildOutput(GpuBuildOutput& out);

}  // namespace yb
