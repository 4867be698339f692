// bvh_build.cpp -- see bvh_build.hpp.  Follows Culling.hs:27-112 and AABBs.hs:10-27,45-53.
#include "bvh_build.hpp"

#include <cstring>
#include <stdexcept>

namespace yb {
namespace {

class Builder {
 public:
  Builder(const std::vector<Box>& bounds, int splitMode, HostBvh& out)
      : bounds_(bounds), splitMode_(splitMode), out_(out) {
    const size_t n = bounds.size();
    cent_.resize(n);
    for (size_t i = 0; i < n; ++i) cent_[i] = boxCentroid(bounds[i]);   // centroid . fst
    idx_.resize(n);
    for (size_t i = 0; i < n; ++i) idx_[i] = (uint32_t)i;
    tmp_.resize(n);
  }

  void run(int maxDepth) {
    out_ = HostBvh();
    if (bounds_.empty()) return;                                        // bvh _ _ [] = const Nothing
    out_.nodes.reserve(2 * bounds_.size());
    buildTree(maxDepth, 0, (uint32_t)idx_.size(), 0);
    out_.order = idx_;                     // stable in-place partitions leave idx_ in DFS leaf order
    flatten();
  }

 private:
  const std::vector<Box>& bounds_;
  int splitMode_;
  HostBvh& out_;
  std::vector<f3> cent_;
  std::vector<uint32_t> idx_, tmp_;

  int32_t newNode(uint32_t level) {
    if (level > out_.depth) out_.depth = level;
    out_.nodes.emplace_back();
    return (int32_t)out_.nodes.size() - 1;
  }

  // multiLeaf (Culling.hs:52-54)
  int32_t multiLeaf(uint32_t lo, uint32_t hi, uint32_t level) {
    int32_t id = newNode(level);
    Box jb = emptyBox();
    for (uint32_t k = lo; k < hi; ++k) jb = joinBox(jb, bounds_[idx_[k]]);   // foldl join empty bbs
    RefNode& n = out_.nodes[id];
    n.kind = kMulti; n.first = lo; n.count = hi - lo; n.left = n.right = -1; n.box = jb;
    return id;
  }

  // Stable partition of idx_[lo,hi) by predicate; returns the split position.
  template <class Pred>
  uint32_t stablePartition(uint32_t lo, uint32_t hi, Pred&& goesLeft) {
    uint32_t l = lo, r = 0;
    for (uint32_t k = lo; k < hi; ++k) {
      uint32_t i = idx_[k];
      if (goesLeft(i)) idx_[l++] = i; else tmp_[r++] = i;
    }
    std::memcpy(idx_.data() + l, tmp_.data(), (size_t)r * sizeof(uint32_t));
    return l;
  }

  // midpointSplit (Culling.hs:56-60)
  uint32_t midpointSplit(uint32_t lo, uint32_t hi, int dim, const Box& bbOfCentroids) {
    const float midpoint = comp(boxCentroid(bbOfCentroids), dim);
    return stablePartition(lo, hi, [&](uint32_t i) { return comp(cent_[i], dim) <= midpoint; });
  }

  // `floor x :: Int` through float2Int#: NaN / out-of-range become INT64_MIN on x86-64.
  static int64_t floorToInt(float x) {
    float f = std::floor(x);
    if (!(f >= -9.2233720368547758e18f && f < 9.2233720368547758e18f)) return INT64_MIN;
    return (int64_t)f;
  }

  // sahSplit (Culling.hs:62-112) including its quirks.
  uint32_t sahSplit(uint32_t lo, uint32_t hi, int dim, const Box& bbOfCentroids) {
    if (!(hi - lo > 4)) return midpointSplit(lo, hi, dim, bbOfCentroids);
    const int nBuckets = 16;
    const float centMin = comp(bbOfCentroids.lo, dim), centMax = comp(bbOfCentroids.hi, dim);
    const float extent = centMax - centMin;
    auto bucketId = [&](uint32_t i) -> int {
      float frac = (comp(cent_[i], dim) - centMin) / extent;
      int64_t b = floorToInt((float)nBuckets * frac);
      if (b < 0) b = 0;
      if (b > nBuckets - 1) b = nBuckets - 1;
      return (int)b;
    };
    int counts[16];
    Box boxes[16];
    for (int b = 0; b < nBuckets; ++b) { counts[b] = 0; boxes[b] = emptyBox(); }
    for (uint32_t k = lo; k < hi; ++k) {
      int b = bucketId(idx_[k]);
      counts[b]++;
      boxes[b] = joinBox(boxes[b], bounds_[idx_[k]]);
    }
    float wrappedArea = 0.0f;
    for (int b = 0; b < nBuckets; ++b) wrappedArea = wrappedArea + boxSurf(boxes[b]);
    // prefix (from the left) and suffix joins replace the reference's per-split refolds; join is
    // associative for non-NaN boxes and each fold starts from `empty` exactly like V.foldl join empty.
    Box pre[16], suf[16];
    int cpre[16], csuf[16];
    Box acc = emptyBox(); int c = 0;
    for (int b = 0; b < nBuckets; ++b) { acc = joinBox(acc, boxes[b]); c += counts[b]; pre[b] = acc; cpre[b] = c; }
    int bestSplit = 0; float bestCost = 0.0f;
    for (int s = 0; s <= nBuckets - 2; ++s) {
      Box b1 = emptyBox(); int c1 = 0;
      for (int b = s + 1; b < nBuckets; ++b) { b1 = joinBox(b1, boxes[b]); c1 += counts[b]; }
      suf[s] = b1; csuf[s] = c1;
      float cost = 0.125f + ((float)cpre[s] * boxSurf(pre[s]) + (float)csuf[s] * boxSurf(suf[s])) / wrappedArea;
      // minimumBy (compare `on` snd): the accumulator is replaced unless acc < next or acc == next
      if (s == 0 || !((bestCost < cost) || (bestCost == cost))) { bestSplit = s; bestCost = cost; }
    }
    return stablePartition(lo, hi, [&](uint32_t i) { return bucketId(i) <= bestSplit; });
  }

  // buildTree (Culling.hs:31-50)
  int32_t buildTree(int maxDepth, uint32_t lo, uint32_t hi, uint32_t level) {
    if (lo == hi) {                                     // buildTree _ [] = (empty, nullCollider)
      int32_t id = newNode(level);
      RefNode& n = out_.nodes[id];
      n.kind = kNull; n.first = lo; n.count = 0; n.left = n.right = -1; n.box = emptyBox();
      return id;
    }
    if (hi - lo == 1) {                                 // buildTree _ [(bb, col)] = (bb, wrapCollider col bb)
      int32_t id = newNode(level);
      RefNode& n = out_.nodes[id];
      n.kind = kSingle; n.first = lo; n.count = 1; n.left = n.right = -1; n.box = bounds_[idx_[lo]];
      return id;
    }
    if (maxDepth == 0) return multiLeaf(lo, hi, level); // buildTree 0 xs = multiLeaf xs

    // bbOfCentroids = boundAllPoints (map centroid bbs)   (AABBs.hs:20-23; hi - lo >= 2 here)
    f3 c0 = cent_[idx_[lo]], c1 = cent_[idx_[lo + 1]];
    Box cb{hmin3(c0, c1), hmax3(c0, c1)};               // fromPoints x1 x2
    for (uint32_t k = lo + 2; k < hi; ++k) {            // foldl includePoint
      f3 c = cent_[idx_[k]];
      cb.lo = hmin3(cb.lo, c);
      cb.hi = hmax3(cb.hi, c);
    }
    const int dim = maxDimension(cb.hi - cb.lo);        // maxExtent
    const uint32_t mid = splitMode_ == 1 ? sahSplit(lo, hi, dim, cb) : midpointSplit(lo, hi, dim, cb);
    if (mid == hi) return multiLeaf(lo, hi, level);     // if null xs2 then multiLeaf xs1

    int32_t id = newNode(level);
    int32_t l = buildTree(maxDepth - 1, lo, mid, level + 1);
    int32_t r = buildTree(maxDepth - 1, mid, hi, level + 1);
    RefNode& n = out_.nodes[id];
    n.kind = kInner; n.first = lo; n.count = hi - lo; n.left = l; n.right = r;
    n.box = joinBox(out_.nodes[l].box, out_.nodes[r].box);   // jointBB = join bb1 bb2
    return id;
  }

  void flatten() {
    const size_t n = out_.nodes.size();
    std::vector<uint32_t> innerIndex(n, 0);
    uint32_t nInner = 0;
    for (size_t i = 0; i < n; ++i)
      if (out_.nodes[i].kind == kInner) innerIndex[i] = nInner++;
    auto refOf = [&](int32_t id) -> uint32_t {
      const RefNode& c = out_.nodes[id];
      switch (c.kind) {
        case kNull: return kRefNull;
        case kInner: return innerIndex[id];
        case kSingle:
          if (c.first >= 0x40000000u) throw std::runtime_error("too many primitives for the leaf encoding");
          return kRefLeafBit | c.first;
        default: {
          uint32_t k = (uint32_t)(out_.multiLeaves.size() / 2);
          out_.multiLeaves.push_back(c.first);
          out_.multiLeaves.push_back(c.count);
          return kRefMultiBits | k;
        }
      }
    };
    if (nInner >= 0x80000000u) throw std::runtime_error("too many BVH nodes");
    out_.flat.resize(nInner);
    for (size_t i = 0; i < n; ++i) {
      const RefNode& nd = out_.nodes[i];
      if (nd.kind != kInner) continue;
      FlatNode& f = out_.flat[innerIndex[i]];
      const Box& lb = out_.nodes[nd.left].box;
      const Box& rb = out_.nodes[nd.right].box;
      f.lxy[0] = lb.lo.x; f.lxy[1] = lb.lo.y; f.lxy[2] = lb.hi.x; f.lxy[3] = lb.hi.y;
      f.rxy[0] = rb.lo.x; f.rxy[1] = rb.lo.y; f.rxy[2] = rb.hi.x; f.rxy[3] = rb.hi.y;
      f.z[0] = lb.lo.z; f.z[1] = lb.hi.z; f.z[2] = rb.lo.z; f.z[3] = rb.hi.z;
      f.left = refOf(nd.left);
      f.right = refOf(nd.right);
      f.pad0 = f.pad1 = 0;
    }
    out_.rootRef = refOf(0);
    out_.rootBox = out_.nodes[0].box;
    out_.maxStack = out_.depth + 1;
  }
};

}  // namespace

void buildReferenceBvh(const std::vector<Box>& bounds, int maxDepth, int splitMode, HostBvh& out) {
  Builder b(bounds, splitMode, out);
  b.run(maxDepth < 0 ? 0 : maxDepth);
}

}  // namespace yb
