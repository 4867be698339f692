// host_scene.hpp -- host-side setup shared by the C ABI: primitive gathering + bounds
// (main.hs:41-53, Shapes.hs:32-33,62-63), camera matrices (Cameras.hs:24-82 with the `linear`
// package's fold order) and the reference's tile arithmetic (Sampling.hs:9-21, main.hs:109-131).
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/yahr_b200.h"
#include "bvh_build.hpp"

namespace yb {

struct HostPrim {
  uint8_t kind;              // 0 sphere, 1 triangle
  f3 a, b, c, n0, n1, n2;    // sphere: a = centre; triangle: p0 p1 p2 + normals
  float radius;
  uint32_t material;
};

// Collects the expanded primitive list in primitive-ID order and their bounds.
// Returns YAHR_OK or an error code with `err` set.
inline int gatherPrimitives(const yahr_scene_desc* d, std::vector<HostPrim>& prims, std::vector<Box>& bounds,
                            std::string& err) {
  if (!d) { err = "scene descriptor is NULL"; return YAHR_ERR_INVALID_ARGUMENT; }
  const uint64_t n = (uint64_t)d->n_spheres + d->n_triangles;
  if (n >= 0x40000000ull) { err = "more than 2^30 primitives"; return YAHR_ERR_INVALID_ARGUMENT; }
  if (d->n_triangles && !(d->tri_p0 && d->tri_p1 && d->tri_p2 && d->tri_n0 && d->tri_n1 && d->tri_n2)) {
    err = "triangle arrays missing"; return YAHR_ERR_INVALID_ARGUMENT;
  }
  if (d->n_spheres && !(d->sph_center && d->sph_radius)) { err = "sphere arrays missing"; return YAHR_ERR_INVALID_ARGUMENT; }
  if (d->n_materials && !d->materials) { err = "materials missing"; return YAHR_ERR_INVALID_ARGUMENT; }
  if (d->n_lights && !d->lights) { err = "lights missing"; return YAHR_ERR_INVALID_ARGUMENT; }
  prims.resize(n);
  bounds.resize(n);
  auto v = [](const float* p, uint32_t k) { return f3{p[3 * (size_t)k], p[3 * (size_t)k + 1], p[3 * (size_t)k + 2]}; };
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t ref = d->prim_order ? d->prim_order[i]
                                       : (i < d->n_spheres ? (uint32_t)i : (0x80000000u | (uint32_t)(i - d->n_spheres)));
    const uint32_t k = ref & 0x7FFFFFFFu;
    HostPrim& p = prims[i];
    p = HostPrim{};
    if (ref >> 31) {
      if (k >= d->n_triangles) { err = "prim_order: triangle index out of range"; return YAHR_ERR_INVALID_ARGUMENT; }
      p.kind = 1;
      p.a = v(d->tri_p0, k); p.b = v(d->tri_p1, k); p.c = v(d->tri_p2, k);
      p.n0 = v(d->tri_n0, k); p.n1 = v(d->tri_n1, k); p.n2 = v(d->tri_n2, k);
      p.material = d->tri_material ? d->tri_material[k] : 0;
      if (!(finite3(p.a) && finite3(p.b) && finite3(p.c))) { err = "non-finite triangle vertex"; return YAHR_ERR_NON_FINITE_INPUT; }
      // boundTriangle = includePoint (fromPoints p0 p1) p2          (Shapes.hs:62-63)
      Box b{hmin3(p.a, p.b), hmax3(p.a, p.b)};
      bounds[i] = Box{hmin3(b.lo, p.c), hmax3(b.hi, p.c)};
    } else {
      if (k >= d->n_spheres) { err = "prim_order: sphere index out of range"; return YAHR_ERR_INVALID_ARGUMENT; }
      p.kind = 0;
      p.a = v(d->sph_center, k);
      p.radius = d->sph_radius[k];
      p.material = d->sph_material ? d->sph_material[k] : 0;
      if (!(finite3(p.a) && std::isfinite(p.radius))) { err = "non-finite sphere"; return YAHR_ERR_NON_FINITE_INPUT; }
      // boundSphere r s = fromPoints (s + vof r) (s - vof r)        (Shapes.hs:32-33)
      const f3 r3{p.radius, p.radius, p.radius};
      const f3 x = p.a + r3, y = p.a - r3;
      bounds[i] = Box{hmin3(x, y), hmax3(x, y)};
    }
    if (p.material >= d->n_materials) {              // the reference dies in `mats ! mId` (main.hs:55)
      err = "material index " + std::to_string(p.material) + " out of range for primitive " + std::to_string(i);
      return YAHR_ERR_UNKNOWN_MATERIAL;
    }
  }
  return YAHR_OK;
}

// ---- camera -------------------------------------------------------------------------------
struct Mat4 { float m[4][4]; };

// linear: f !*! g = fmap (\f' -> foldl' (^+^) zero $ liftI2 (*^) f' g) f   (left fold from zero)
inline Mat4 matMul(const Mat4& f, const Mat4& g) {
  Mat4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float acc = 0.0f;
      for (int k = 0; k < 4; ++k) acc = acc + f.m[i][k] * g.m[k][j];
      r.m[i][j] = acc;
    }
  return r;
}

struct CameraSetup {
  Mat4 tf, vtf;       // cameraToWorld, tf !*! rasterToCamera
  float origin[3];    // transformPoint tf (0,0,0)
  int width, height;  // floor imW, floor imH  (main.hs:122-123)
};

inline int setupCamera(const yahr_camera* c, CameraSetup& out, std::string& err) {
  if (!c) { err = "camera is NULL"; return YAHR_ERR_INVALID_ARGUMENT; }
  const f3 look{c->lookDir[0], c->lookDir[1], c->lookDir[2]}, upDir{c->upDir[0], c->upDir[1], c->upDir[2]};
  const f3 pos{c->position[0], c->position[1], c->position[2]};
  // cameraToWorld (Cameras.hs:59-65)
  const f3 forward = normalize(look);
  const f3 left = normalize(cross(forward, upDir));
  const f3 up = cross(left, forward);
  const f3 nl = -left;
  const Mat4 dirTf{{{nl.x, up.x, forward.x, 0}, {nl.y, up.y, forward.y, 0}, {nl.z, up.z, forward.z, 0}, {0, 0, 0, 1}}};
  const Mat4 tr{{{1, 0, 0, pos.x}, {0, 1, 0, pos.y}, {0, 0, 1, pos.z}, {0, 0, 0, 1}}};
  out.tf = matMul(tr, dirTf);
  // rasterToCamera (Cameras.hs:68-75)
  const float frameW = c->imW / c->imH, frameH = 1.0f;
  const Mat4 r2c{{{frameW / c->imW, 0, 0, -(frameW / 2.0f)}, {0, -(frameH / c->imH), 0, frameH / 2.0f}, {0, 0, 1, 0}, {0, 0, 0, 1}}};
  out.vtf = matMul(out.tf, r2c);
  // origin = transformPoint tf (Vec3 0 0 0): m !* V4 0 0 0 1, then normalizePoint
  float tv[4];
  const float p[4] = {0.0f, 0.0f, 0.0f, 1.0f};
  for (int i = 0; i < 4; ++i) {
    float acc = 0.0f;
    for (int k = 0; k < 4; ++k) acc = acc + out.tf.m[i][k] * p[k];
    tv[i] = acc;
  }
  const float iw = 1.0f / tv[3];
  for (int i = 0; i < 3; ++i) out.origin[i] = iw * tv[i];
  out.width = (int)std::floor(c->imW);
  out.height = (int)std::floor(c->imH);
  if (!(out.width > 0 && out.height > 0) || (int64_t)out.width * out.height > 0x7FFFFFFF / 3) {
    err = "image size out of range"; return YAHR_ERR_INVALID_ARGUMENT;
  }
  // The kernels skip normalizePoint's (1/w): require the homogeneous row to be exactly (0,0,0,1)
  // and every entry finite, which holds for any finite, non-degenerate camera.
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (!std::isfinite(out.vtf.m[i][j]) || !std::isfinite(out.tf.m[i][j])) {
        err = "degenerate camera (non-finite matrix)"; return YAHR_ERR_INVALID_ARGUMENT;
      }
  if (!(out.vtf.m[3][0] == 0 && out.vtf.m[3][1] == 0 && out.vtf.m[3][2] == 0 && out.vtf.m[3][3] == 1) ||
      !std::isfinite(c->focalLength)) {
    err = "degenerate camera"; return YAHR_ERR_INVALID_ARGUMENT;
  }
  return YAHR_OK;
}

// ---- tiles --------------------------------------------------------------------------------
struct TileWindow { int32_t x0, y0, x1, y1; };

// batchWindow (Sampling.hs:9-21); the divisions default to Double in the reference.
inline TileWindow batchWindow(int64_t w, int64_t h, int64_t num, int64_t count) {
  int64_t nx = count, ny = 1;
  while (nx % 2 == 0 && 2 * w * ny < h * nx) { nx /= 2; ny *= 2; }
  const int64_t j = num / nx, i = num % nx;
  TileWindow t;
  t.x0 = (int32_t)std::floor((double)i / (double)nx * (double)w);
  t.y0 = (int32_t)std::floor((double)j / (double)ny * (double)h);
  t.x1 = (int32_t)std::floor(((double)i + 1.0) / (double)nx * (double)w);
  t.y1 = (int32_t)std::floor(((double)j + 1.0) / (double)ny * (double)h);
  return t;
}
// roundUpPow2 (main.hs:109-110) and nBatches (main.hs:128-130):
//   roundUpPow2 $ max (32 * numThreads) width * height `div` (16 * 16)
//   == roundUpPow2 (((max (32 * numThreads) width) * height) `div` 256)
inline int64_t numBatches(int64_t numThreads, int64_t width, int64_t height) {
  const int64_t m = 32 * numThreads > width ? 32 * numThreads : width;
  const int64_t x = (m * height) / 256;
  if (x < 1) return 1;   // the reference dies here ("Negative exponent"); images this small get one tile
  const double e = std::ceil(std::log((double)x) / std::log(2.0));   // logBase 2 in Double
  int64_t r = 1;
  for (int64_t k = 0; k < (int64_t)e; ++k) r *= 2;
  return r;
}

}  // namespace yb
