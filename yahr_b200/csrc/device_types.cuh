// device_types.cuh -- data layout in HBM and the kernel parameter block (shared by host and device code).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace yb {

// Flattened scene (see DESIGN.md "Data layout in HBM").
//   nodes     : 4 x float4 per inner node                                                      64 B
//                 (L.lo.x L.lo.y L.hi.x L.hi.y) (R.lo.x R.lo.y R.hi.x R.hi.y)
//                 (L.lo.z L.hi.z R.lo.z R.hi.z) (left ref, right ref, -, -)
//               components are paired so that the slab arithmetic runs as packed f32x2 ops
//   prims     : 3 x float4 per primitive in left-first DFS leaf order                          48 B
//                 triangle: A = (p0.xyz, meta)  B = (e1.xyz, primitive ID)  C = (e2.xyz, 0)
//                 sphere  : A = (c.xyz,  meta)  B = (r, 0, 0, primitive ID) C unused
//                 meta = (material << 1) | kind   (kind 0 sphere, 1 triangle), as raw bits
//               e1 = p1 - p0 and e2 = p2 - p0 are the reference's own first two operations
//               (Shapes.hs:38-39) done once on the host in binary32.
//   normals   : 3 x float4 per primitive (n0, n1, n2), same order; only read for triangle
//               candidates that pass the barycentric / t tests                                 48 B
//   wide      : 8 x float4 per 4-wide node (collapse of the binary tree, wide_bvh.cu)           128 B
//                 child k = 0..3: (lo.x lo.y hi.x hi.y); then (z of child 0, 1) (z of child 2, 3) as
//                 (lo.z hi.z lo.z hi.z); then the four child refs; then (child count, binary root, -, -).
//               Inner refs index `wide`; leaf refs are the binary tree's.  Empty slots hold the box
//               (+inf, -inf) and the null ref.  Used by rays whose 1/u components are all finite.
//   multiLeaves : (first, count) per multi-leaf
//   materials : 2 x float4 = (diffuse.rgb, shininess), (specular.rgb, 0)
//   lights    : 2 x float4 = (position.xyz, 0), (spectrum.rgb, 0)
//   areaLights: 5 x float4 = (corner.xyz, samples as bits), (edge1.xyz, 0), (edge2.xyz, 0),
//               (normal.xyz, 0), (flux.rgb, 0)      -- extension, see include/yahr_b200.h
//   Light SLOTS number the point lights first, then every sample of every area light (nSlots in all).
struct DeviceScene {
  const float4* nodes;
  const float4* wide;
  const float4* cwide;      // compressed copy of `wide` (64 B per node, wide_bvh.cu) or NULL
  const float4* leafBox;    // exact boxes of the single-primitive leaves, by DFS position (compressed walk)
  const float4* multiBox;   // exact boxes of the multi-leaves
  const float4* prims;
  const float4* normals;
  const uint2* multiLeaves;
  const float4* materials;
  const float4* lights;
  const float4* areaLights;
  uint32_t nLights;
  uint32_t nAreaLights;
  uint32_t nSlots;
  uint32_t rootRef;
  float rootLo[3], rootHi[3];
};

static const uint32_t kDevRefNull = 0xFFFFFFFFu;
static const uint32_t kDevRefLeafBit = 0x80000000u;
static const uint32_t kDevRefMultiBits = 0xC0000000u;
static const int kWideWidth = 4;          // children per wide node
static const int kWideNodeVec = 8;        // float4 per wide node

struct RenderParams {
  DeviceScene sc;
  float vtf[12];          // rows 0..2 of vtf = tf !*! rasterToCamera (Cameras.hs:81)
  float origin[3];        // transformPoint tf (0,0,0)              (Cameras.hs:84)
  float focal;
  int32_t width, height;
  int32_t depth, spp;
  uint64_t seed;
  int32_t traversal;
  const int4* tiles;      // windows (x0, y0, x1, y1) of the tiles rendered by this call
  uint32_t nTiles;
  float* rgb;             // W*H*3, may be peer memory
  uint32_t* primid;       // W*H or nullptr
  unsigned long long* counters;  // [0] primary, [1] shadow, [2] secondary
};

// Parameter block of the wavefront kernel set (wavefront.cu).  "item" = one pixel of this call's
// tiles, numbered tile-major (tile order, then u-major inside the tile, Sampling.hs:6).
struct WavefrontParams {
  RenderParams base;
  const uint32_t* tileStart;  // nTiles + 1 prefix sums of the tile pixel counts
  uint32_t nItems;            // items of this launch (a range of tiles)
  uint32_t itemBase;          // tileStart value of the first tile of this launch
  const uint32_t* itemPixels; // per item of the tile set (index item + itemBase): u | v << 16, or NULL (computed)
  uint32_t sample;            // first sample index of this launch
  uint32_t samplesPerLaunch;  // spp > 1: one launch traces this many samples of every pixel (work = sample-major)
  uint32_t itemsPadded;       // nItems rounded up to a multiple of 32: work items per sample
  uint32_t framePixels;       // width * height: pixels between the per-sample frames of sampleOut
  uint32_t dense;             // three or more light slots: shadow entry = item * nSlots + slot, k_wf_resolve sums in slot order
  uint32_t twoSlot;           // exactly two light slots: compacted queue; the pixel gets its base value from the shading
                              // step and every unoccluded probe ADDS its contribution with a float atomic -- with two
                              // addends on top of +0 the sum does not depend on their order ((0 + a) + b = (0 + b) + a)
  uint32_t leafThreshold;     // leaf parking: run the leaf code once this many lanes hold a leaf
  uint32_t blocksPerSM;       // tuning: persistent CTAs per SM (0 = as many as fit)
  uint32_t capRegisters;      // tuning: primary kernel compiled for 8 CTAs/SM (<= 64 registers)
  uint32_t packed;            // octant-specialised packed-f32x2 node step when a warp shares an octant
  uint32_t wideTree;          // traverse the 4-wide collapse of the tree (reference order only)
  uint32_t compressed;        // ... through its compressed 64-byte nodes (conservative inner boxes, exact leaf boxes)
  uint32_t split;             // three-kernel set: k_wf_trace (closest-hit walk only) -> k_wf_shade -> k_wf_shadow
  uint2* hits;                // split: per (sample of the launch, item) the hit's DFS position and t
  uint32_t leafRun;           // wide walk: consecutive pending leaves of a lane are tested in one leaf phase
  uint32_t fused;             // one light slot: k_wf_fused (each warp walks its own batch's shadow probes) instead of
                              // primary + shadow; bit 1: the 7-CTAs-per-SM build (72 registers)
  float4* q0;                 // shadow probes: (origin.xyz, tMax)
  float4* q1;                 //                (direction.xyz, pixel index bits)
  float4* q2;                 //                (contribution.rgb, -)
  unsigned char* visibility;  // dense mode: per slot 1 = unoccluded
  uint32_t* bandStat;         // host-buffer entry: += shadow probes of this launch (cost of the band, for scheduling) or NULL
  uint32_t* work;             // [0] primary work counter, [1] shadow work counter / chunks claimed, [2] queue length,
                              // [3] probes, [4] secondary rays, [5] k_wf_persist: primary batches whose probes are committed
  float* sampleOut;           // where this pass writes its radiance (the frame, or sampleBuf for spp > 1)
  float* sampleBuf;           // W*H*3 scratch (spp > 1)
  float* accum;               // W*H*3 running sum (spp > 1)
  // Streamed host output (host-buffer entry, one light slot, 1 spp; NULL / 0 otherwise).  The kernels count the pixels
  // they have finalised per ROW of the tile grid; whoever completes a row publishes it in mapped host memory, and the
  // host thread inside yahr_b200_render starts that row's device-to-host copy while the rest of the frame is traced.
  const unsigned short* rowOfV;   // image row v -> index of its tile row in the tile set
  const uint32_t* rowItems;       // pixels per tile row
  uint32_t* rowDone;              // finalised pixels per tile row (zeroed before the launch)
  volatile uint32_t* rowFlags;    // mapped host memory: rowFlags[row] = rowSeq once the row is complete in device memory
  uint32_t rowSeq;
  // k_wf_persist (one launch per frame): the probes go through a 256-entry ring per CTA in q0 / q2 (32 B per probe)
  uint32_t persist;               // 1: one light slot, depth 1, 4-wide tree -> k_wf_persist instead of primary + shadow
  uint32_t stackShared;           // traversal-stack entries per lane kept in shared memory (0, 8 or 12)
  unsigned long long* workStats;  // counting build only: 2 x 8 work counters (closest-hit walks, any-hit walks); else NULL
  unsigned char* rgb8;            // 8-bit host-buffer entry through the streamed rows: the per-batch kernels store the
                                  // quantised pixel (main.hs:142) here INSTEAD of the float frame; NULL otherwise
};

}  // namespace yb
