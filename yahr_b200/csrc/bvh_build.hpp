// bvh_build.hpp -- host construction of the reference's BVH as DATA.
//
// In the reference the BVH is not a data structure: `cull` returns a closure built from nested
// wrapCollider closures (Culling.hs:21-54).  To traverse it on a GPU the library re-runs the same
// top-down build (same split dimension, same midpoint / 16-bucket SAH decisions, same stable
// partition, same depth cap) and emits
//   * the reference tree in pre-order (for inspection and the parity tests), and
//   * a flattened device layout: 64-byte inner nodes that hold BOTH children's boxes (a box test
//     "on entry to a child" is the same arithmetic wherever the box is stored) and the primitives
//     permuted into left-first DFS leaf order.
#pragma once
#include <cstdint>
#include <vector>

#include "host_math.hpp"

namespace yb {

enum RefKind : uint32_t { kNull = 0, kSingle = 1, kInner = 2, kMulti = 3 };

struct RefNode {
  uint32_t kind;
  uint32_t first, count;   // leaf range in DFS order
  int32_t left, right;     // inner
  Box box;                 // the box given to wrapCollider
};

// Child reference encoding of the flattened layout.
static const uint32_t kRefNull = 0xFFFFFFFFu;
static const uint32_t kRefLeafBit = 0x80000000u;   // single-primitive leaf: kRefLeafBit | dfs position
static const uint32_t kRefMultiBits = 0xC0000000u; // multi-leaf: kRefMultiBits | index into multiLeaves

struct FlatNode {          // 64 bytes = 4 x float4, components paired for packed f32x2 slab math
  float lxy[4];            // left child:  lo.x lo.y hi.x hi.y
  float rxy[4];            // right child: lo.x lo.y hi.x hi.y
  float z[4];              // left lo.z, left hi.z, right lo.z, right hi.z
  uint32_t left, right;    // child references
  uint32_t pad0, pad1;
};
static_assert(sizeof(FlatNode) == 64, "FlatNode must be 64 bytes");

struct HostBvh {
  std::vector<RefNode> nodes;          // pre-order; nodes[0] is the root when non-empty
  std::vector<uint32_t> order;         // order[i] = primitive ID at DFS leaf position i
  uint32_t depth = 0;                  // deepest level reached (root = 0)
  // flattened
  std::vector<FlatNode> flat;
  std::vector<uint32_t> multiLeaves;   // pairs (first, count)
  uint32_t rootRef = kRefNull;
  Box rootBox = emptyBox();
  uint32_t maxStack = 0;               // pending right children a left-first walk can hold
};

// bounds[i] = boundSceneObject of primitive i (main.hs:51-53).  splitMode 0 Midpoint, 1 SAH.
void buildReferenceBvh(const std::vector<Box>& bounds, int maxDepth, int splitMode, HostBvh& out);

}  // namespace yb
