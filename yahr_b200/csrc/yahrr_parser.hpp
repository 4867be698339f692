// yahrr_parser.hpp -- host-side mirror of the reference's scene description (Scene.hs:15-58) and of
// `Scene.expand` (Scene.hs:61-86).
//
// The `.yahrr` format IS Haskell's derived `Read` syntax of `Scene` (main.hs:117).  No GHC exists in the
// build image, so the C++ host reads that grammar itself: record syntax with the fields in declaration
// order, positional constructors, Just/Nothing, lists, 3-tuples, string literals, ints / decimals /
// exponents, negative numbers with or without parentheses (both occur: scene.yahrr:6,35;
// compat/yahr.py:142 prints bare `-0.100000`), redundant parentheses around any value.
// Leniency (documented): a TriangleMesh may omit `triangleMeshNormals` / `triangleMeshSmooth`
// (the repo's own, stale scene.yahrr does; they default to Nothing).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/yahr_b200.h"

namespace yb {

struct SceneMaterial {      // Scene.BlinnPhongMaterial (Scene.hs:45-50)
  std::string id;
  float ambient[3], diffuse[3], specular[3], shininess;
};

// A parsed scene with `objects >>= expand` already applied (main.hs:44): arrays in the layout of
// yahr_scene_desc.  `desc()` points into this object; keep it alive while the descriptor is used.
struct LoadedScene {
  int recursionDepth = 1;                 // Integrators.WhittedIntegrator (Integrators.hs:18-20)
  int bvhMaxDepth = 16, splitMode = 0;    // Culling.BVH Int SplitMode (Culling.hs:18-19)
  yahr_camera camera{};                   // Cameras.Camera (Cameras.hs:54-56)
  std::vector<SceneMaterial> materials;
  std::vector<float> lights;              // 6 per light
  std::vector<float> triP0, triP1, triP2, triN0, triN1, triN2;
  std::vector<uint32_t> triMaterial;
  std::vector<float> sphCenter, sphRadius;
  std::vector<uint32_t> sphMaterial;
  std::vector<uint32_t> primOrder;        // (kind << 31) | index, in `objects >>= expand` order
  std::vector<float> materials7;          // diffuse, specular, shininess per material
  yahr_scene_desc desc() const;
};

// Parses `.yahrr` text and expands the objects.  Returns YAHR_OK, YAHR_ERR_PARSE (the reference:
// "Prelude.read: no parse") or YAHR_ERR_UNKNOWN_MATERIAL (the reference: Map.! error, main.hs:55).
int loadYahrr(const std::string& text, LoadedScene& out, std::string& err);

// JuicyPixels' float -> 8-bit conversion used by savePngImage (ImageRGBF) (main.hs:142):
// truncate (255 * max 0 (min 1 x)), no gamma; GHC's min/max make NaN -> 0.
inline uint8_t quantize8(float x) {
  float m = (1.0f <= x) ? 1.0f : x;      // min 1 x
  float c = (0.0f <= m) ? m : 0.0f;      // max 0 m
  return (uint8_t)(int)(255.0f * c);
}

// Writes an 8-bit RGB PNG (stored deflate blocks; no external dependency).
int writePngRgb8(const std::string& path, const uint8_t* rgb, int width, int height, std::string& err);

}  // namespace yb
