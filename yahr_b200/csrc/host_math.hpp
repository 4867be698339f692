// host_math.hpp -- binary32 host-side vector helpers of libyahr_b200 (setup code: camera matrices,
// bounds, BVH construction).  Expression order follows the reference's Vectors.hs so that the
// values handed to the kernels are the ones the Haskell program would compute.  Compiled with
// -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>

namespace yb {

struct f3 { float x, y, z; };

inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
inline f3 operator*(float s, f3 v) { return {s * v.x, s * v.y, s * v.z}; }          // (@*) Vectors.hs:42-44
inline float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }        // (.*) Vectors.hs:32-34
inline f3 cross(f3 a, f3 b) {                                                       // Vectors.hs:49-53
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline f3 normalize(f3 v) { return (1.0f / std::sqrt(dot(v, v))) * v; }             // norm Vectors.hs:46-47
inline float comp(f3 v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }

// GHC `Ord Float` class-default min/max (SURVEY.md note N): selects, not fminf/fmaxf.
inline float hmin(float x, float y) { return x <= y ? x : y; }
inline float hmax(float x, float y) { return x <= y ? y : x; }
inline f3 hmin3(f3 a, f3 b) { return {hmin(a.x, b.x), hmin(a.y, b.y), hmin(a.z, b.z)}; }
inline f3 hmax3(f3 a, f3 b) { return {hmax(a.x, b.x), hmax(a.y, b.y), hmax(a.z, b.z)}; }

struct Box { f3 lo, hi; };
inline Box emptyBox() { const float inf = INFINITY; return {{inf, inf, inf}, {-inf, -inf, -inf}}; }  // AABBs.hs:10-11
inline Box joinBox(const Box& a, const Box& b) { return {hmin3(a.lo, b.lo), hmax3(a.hi, b.hi)}; }    // AABBs.hs:25-27
inline f3 boxCentroid(const Box& b) { return 0.5f * b.lo + 0.5f * b.hi; }                            // AABBs.hs:48-49
inline float boxSurf(const Box& b) {                                                                 // AABBs.hs:51-53
  f3 d = b.hi - b.lo;
  return 2.0f * ((d.x * d.y + d.x * d.z) + d.y * d.z);
}
// maxDimension (Vectors.hs:62-66): strict >, ties fall through X -> Y -> Z
inline int maxDimension(f3 v) { return (v.x > v.y && v.x > v.z) ? 0 : (v.y > v.z ? 1 : 2); }

inline bool finite3(f3 v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }

}  // namespace yb
