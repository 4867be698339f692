// yahrr_parser.cpp -- see yahrr_parser.hpp.  Follows Scene.hs:15-86, Cameras.hs:54-56, Lights.hs:7,
// Integrators.hs:18-20, Culling.hs:18-19 for the grammar and Scene.hs:61-86 for `expand`.
#include "yahrr_parser.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>

#include "host_math.hpp"

namespace yb {
namespace {

// ---- lexer ---------------------------------------------------------------------------------
enum TokKind { kIdent, kNumber, kString, kPunct, kEnd };
struct Token { TokKind kind; std::string text; size_t pos; };

bool lex(const std::string& s, std::vector<Token>& out, std::string& err) {
  size_t i = 0, n = s.size();
  while (i < n) {
    const char c = s[i];
    if (c == ' ' || c == '\t' || c == '\n' || c == '\r') { ++i; continue; }
    if (c == '-' && i + 1 < n && s[i + 1] == '-') {               // Haskell line comment
      while (i < n && s[i] != '\n') ++i;
      continue;
    }
    if (std::isalpha((unsigned char)c) || c == '_') {
      size_t j = i;
      while (j < n && (std::isalnum((unsigned char)s[j]) || s[j] == '_' || s[j] == '\'')) ++j;
      out.push_back({kIdent, s.substr(i, j - i), i});
      i = j;
      continue;
    }
    if (std::isdigit((unsigned char)c)) {
      size_t j = i;
      while (j < n && std::isdigit((unsigned char)s[j])) ++j;
      if (j + 1 < n && s[j] == '.' && std::isdigit((unsigned char)s[j + 1])) {
        ++j;
        while (j < n && std::isdigit((unsigned char)s[j])) ++j;
      }
      if (j < n && (s[j] == 'e' || s[j] == 'E')) {
        size_t k = j + 1;
        if (k < n && (s[k] == '+' || s[k] == '-')) ++k;
        if (k < n && std::isdigit((unsigned char)s[k])) {
          while (k < n && std::isdigit((unsigned char)s[k])) ++k;
          j = k;
        }
      }
      out.push_back({kNumber, s.substr(i, j - i), i});
      i = j;
      continue;
    }
    if (c == '"') {
      std::string v;
      size_t j = i + 1;
      bool closed = false;
      while (j < n) {
        if (s[j] == '"') { closed = true; ++j; break; }
        if (s[j] == '\\' && j + 1 < n) {
          const char e = s[j + 1];
          if (e == 'n') v += '\n';
          else if (e == 't') v += '\t';
          else v += e;                                               // \\ and \" (compat/yahr.py:145-146)
          j += 2;
          continue;
        }
        v += s[j++];
      }
      if (!closed) { err = "unterminated string literal at offset " + std::to_string(i); return false; }
      out.push_back({kString, v, i});
      i = j;
      continue;
    }
    if (std::strchr("{}[](),=-", c)) {
      out.push_back({kPunct, std::string(1, c), i});
      ++i;
      continue;
    }
    err = std::string("unexpected character '") + c + "' at offset " + std::to_string(i);
    return false;
  }
  out.push_back({kEnd, "", n});
  return true;
}

// ---- AST of Scene.SceneObject (Scene.hs:15-42) ------------------------------------------------
struct Obj {
  enum Kind { Sphere, Triangle, Mesh, Subsampled } kind = Sphere;
  f3 position{}, p0{}, p1{}, p2{}, n0{}, n1{}, n2{};
  float radius = 0, subsampleSize = 0;
  std::string materialId;
  std::vector<f3> points, normals;
  bool hasNormals = false, hasSmooth = false;
  std::vector<int64_t> tris;      // 3 per triangle
  std::vector<uint8_t> smooth;
  std::vector<Obj> children;
};

struct Parser {
  std::vector<Token> t;
  size_t i = 0;
  std::string err;
  size_t errPos = 0;

  bool failAt(const std::string& m) {
    if (err.empty() || t[i].pos >= errPos) { err = m + " at offset " + std::to_string(t[i].pos); errPos = t[i].pos; }
    return false;
  }
  bool isP(char c) const { return t[i].kind == kPunct && t[i].text[0] == c; }
  bool eat(char c) { if (isP(c)) { ++i; return true; } return false; }
  bool expect(char c) { return eat(c) || failAt(std::string("expected '") + c + "'"); }
  bool ident(const char* name) {
    if (t[i].kind == kIdent && t[i].text == name) { ++i; return true; }
    return failAt(std::string("expected ") + name);
  }
  bool peekIdent(const char* name) const { return t[i].kind == kIdent && t[i].text == name; }

  // value wrapped in any number of redundant parentheses
  template <class F>
  bool parens(F&& f) {
    int k = 0;
    while (isP('(')) { ++i; ++k; }
    if (!f()) return false;
    while (k-- > 0) if (!expect(')')) return false;
    return true;
  }

  bool number(double& v, std::string* text = nullptr) {       // optional '-' then a numeric literal
    return parens([&] {
      bool neg = eat('-');
      if (t[i].kind != kNumber) return failAt("expected a number");
      if (text) *text = (neg ? "-" : "") + t[i].text;
      v = std::strtod(t[i].text.c_str(), nullptr);
      if (neg) v = -v;
      ++i;
      return true;
    });
  }
  bool floatv(float& v) {                                      // `read :: Float` = correctly rounded decimal
    std::string text;
    double d;
    if (!number(d, &text)) return false;
    v = std::strtof(text.c_str(), nullptr);
    return true;
  }
  bool intv(int64_t& v) {
    std::string text;
    double d;
    if (!number(d, &text)) return false;
    if (text.find_first_of(".eE") != std::string::npos) return failAt("expected an integer");
    v = std::strtoll(text.c_str(), nullptr, 10);
    return true;
  }
  bool vec3(f3& v) {                                           // Vec3 Float Float Float (Vectors.hs:5)
    return parens([&] { return ident("Vec3") && floatv(v.x) && floatv(v.y) && floatv(v.z); });
  }
  bool str(std::string& s) {
    return parens([&] {
      if (t[i].kind != kString) return failAt("expected a string");
      s = t[i].text; ++i;
      return true;
    });
  }
  bool boolv(bool& b) {
    return parens([&] {
      if (peekIdent("True")) { b = true; ++i; return true; }
      if (peekIdent("False")) { b = false; ++i; return true; }
      return failAt("expected True or False");
    });
  }
  template <class F>
  bool list(F&& item) {                                        // [a, b, c]
    return parens([&] {
      if (!expect('[')) return false;
      if (eat(']')) return true;
      for (;;) {
        if (!item()) return false;
        if (eat(',')) continue;
        return expect(']');
      }
    });
  }
  bool field(const char* name) { return ident(name) && expect('='); }
  bool tuple3(int64_t& a, int64_t& b, int64_t& c) {
    // '(' a ',' b ',' c ')' possibly inside redundant parentheses: try the tuple first, then peel one
    const size_t save = i;
    if (eat('(')) {
      if (intv(a) && expect(',') && intv(b) && expect(',') && intv(c) && expect(')')) return true;
      i = save;
      if (eat('(') && tuple3(a, b, c) && expect(')')) return true;
    }
    i = save;
    return failAt("expected a triple (i, j, k)");
  }

  bool object(Obj& o);
  bool scene(LoadedScene& out, std::vector<Obj>& objs);
};

bool Parser::object(Obj& o) {
  return parens([&] {
    if (peekIdent("Sphere")) {                                  // Scene.hs:16-19
      ++i; o.kind = Obj::Sphere;
      return expect('{') && field("position") && vec3(o.position) && expect(',') && field("radius") &&
             floatv(o.radius) && expect(',') && field("materialId") && str(o.materialId) && expect('}');
    }
    if (peekIdent("Triangle")) {                                // Scene.hs:20-28
      ++i; o.kind = Obj::Triangle;
      return expect('{') && field("p0") && vec3(o.p0) && expect(',') && field("p1") && vec3(o.p1) && expect(',') &&
             field("p2") && vec3(o.p2) && expect(',') && field("n0") && vec3(o.n0) && expect(',') && field("n1") &&
             vec3(o.n1) && expect(',') && field("n2") && vec3(o.n2) && expect(',') && field("materialId") &&
             str(o.materialId) && expect('}');
    }
    if (peekIdent("TriangleMesh")) {                            // Scene.hs:29-36
      ++i; o.kind = Obj::Mesh;
      if (!(expect('{') && field("triangleMeshPoints") &&
            list([&] { f3 v; if (!vec3(v)) return false; o.points.push_back(v); return true; }) && expect(',')))
        return false;
      if (peekIdent("triangleMeshNormals")) {                   // Maybe [Vec3]
        if (!field("triangleMeshNormals")) return false;
        if (!parens([&] {
              if (peekIdent("Nothing")) { ++i; return true; }
              if (!ident("Just")) return false;
              o.hasNormals = true;
              return list([&] { f3 v; if (!vec3(v)) return false; o.normals.push_back(v); return true; });
            }) || !expect(','))
          return false;
      }
      if (!(field("triangleMeshTriangles") &&
            list([&] {
              int64_t a, b, c;
              if (!tuple3(a, b, c)) return false;
              o.tris.push_back(a); o.tris.push_back(b); o.tris.push_back(c);
              return true;
            }) && expect(',')))
        return false;
      if (peekIdent("triangleMeshSmooth")) {                    // Maybe [Bool]
        if (!field("triangleMeshSmooth")) return false;
        if (!parens([&] {
              if (peekIdent("Nothing")) { ++i; return true; }
              if (!ident("Just")) return false;
              o.hasSmooth = true;
              return list([&] { bool b = false; if (!boolv(b)) return false; o.smooth.push_back(b ? 1 : 0); return true; });
            }) || !expect(','))
          return false;
      }
      return field("materialId") && str(o.materialId) && expect('}');
    }
    if (peekIdent("Subsampled")) {                              // Scene.hs:37-40
      ++i; o.kind = Obj::Subsampled;
      return expect('{') && field("subsampleSize") && floatv(o.subsampleSize) && expect(',') &&
             field("subsampledObjects") &&
             list([&] { o.children.emplace_back(); return object(o.children.back()); }) && expect('}');
    }
    return failAt("expected Sphere, Triangle, TriangleMesh or Subsampled");
  });
}

bool Parser::scene(LoadedScene& out, std::vector<Obj>& objs) {
  // Scene { integrator, cullingMode, camera, materials, lights, objects }   (Scene.hs:52-58)
  return parens([&] {
    if (!(ident("Scene") && expect('{') && field("integrator"))) return false;
    if (!parens([&] {                                           // Integrators.hs:18-20
          int64_t d;
          if (!(ident("WhittedIntegrator") && expect('{') && field("recursionDepth") && intv(d) && expect('}')))
            return false;
          out.recursionDepth = (int)d;
          return true;
        }) || !expect(','))
      return false;
    if (!field("cullingMode")) return false;
    if (!parens([&] {                                           // BVH Int SplitMode (Culling.hs:18-19)
          int64_t d;
          if (!(ident("BVH") && intv(d))) return false;
          out.bvhMaxDepth = (int)d;
          return parens([&] {
            if (peekIdent("Midpoint")) { ++i; out.splitMode = YAHR_SPLIT_MIDPOINT; return true; }
            if (peekIdent("SurfaceAreaHeuristic")) { ++i; out.splitMode = YAHR_SPLIT_SAH; return true; }
            return failAt("expected Midpoint or SurfaceAreaHeuristic");
          });
        }) || !expect(','))
      return false;
    if (!field("camera")) return false;
    if (!parens([&] {                                           // Cameras.hs:54-56
          f3 look, up, pos;
          if (!(ident("Camera") && expect('{') && field("imW") && floatv(out.camera.imW) && expect(',') &&
                field("imH") && floatv(out.camera.imH) && expect(',') && field("focalLength") &&
                floatv(out.camera.focalLength) && expect(',') && field("lookDir") && vec3(look) && expect(',') &&
                field("upDir") && vec3(up) && expect(',') && field("position") && vec3(pos) && expect('}')))
            return false;
          out.camera.lookDir[0] = look.x; out.camera.lookDir[1] = look.y; out.camera.lookDir[2] = look.z;
          out.camera.upDir[0] = up.x; out.camera.upDir[1] = up.y; out.camera.upDir[2] = up.z;
          out.camera.position[0] = pos.x; out.camera.position[1] = pos.y; out.camera.position[2] = pos.z;
          return true;
        }) || !expect(','))
      return false;
    if (!(field("materials") &&
          list([&] {                                            // Scene.hs:45-50
            return parens([&] {
              SceneMaterial m{};
              f3 a, d, s;
              if (!(ident("BlinnPhongMaterial") && expect('{') && field("id") && str(m.id) && expect(',') &&
                    field("ambient") && vec3(a) && expect(',') && field("diffuse") && vec3(d) && expect(',') &&
                    field("specular") && vec3(s) && expect(',') && field("shininess") && floatv(m.shininess) &&
                    expect('}')))
                return false;
              m.ambient[0] = a.x; m.ambient[1] = a.y; m.ambient[2] = a.z;
              m.diffuse[0] = d.x; m.diffuse[1] = d.y; m.diffuse[2] = d.z;
              m.specular[0] = s.x; m.specular[1] = s.y; m.specular[2] = s.z;
              out.materials.push_back(m);
              return true;
            });
          }) && expect(',')))
      return false;
    if (!(field("lights") &&
          list([&] {                                            // PointLight Vec3 Spectrum (Lights.hs:7)
            return parens([&] {
              f3 p, s;
              if (!(ident("PointLight") && vec3(p) && vec3(s))) return false;
              const float v[6] = {p.x, p.y, p.z, s.x, s.y, s.z};
              out.lights.insert(out.lights.end(), v, v + 6);
              return true;
            });
          }) && expect(',')))
      return false;
    if (!(field("objects") && list([&] { objs.emplace_back(); return object(objs.back()); }) && expect('}')))
      return false;
    return true;
  });
}

// ---- expand (Scene.hs:61-86) --------------------------------------------------------------------
struct Expanded {      // Sphere or Triangle after expansion
  bool triangle;
  f3 a, b, c, n0, n1, n2;
  float radius;
  const std::string* materialId;
};

bool expandObject(const Obj& o, std::vector<Expanded>& out, std::string& err) {
  switch (o.kind) {
    case Obj::Mesh: {
      // zipWith triangleOfPoints triangleMeshTriangles smooth ; smooth = fromMaybe (repeat False) ...
      size_t nt = o.tris.size() / 3;
      if (o.hasSmooth && o.smooth.size() < nt) nt = o.smooth.size();          // zipWith stops at the shorter list
      for (size_t k = 0; k < nt; ++k) {
        const int64_t i0 = o.tris[3 * k], i1 = o.tris[3 * k + 1], i2 = o.tris[3 * k + 2];
        const int64_t np = (int64_t)o.points.size();
        if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= np || i1 >= np || i2 >= np) {
          err = "TriangleMesh: point index out of range";                       // the reference: A.! error
          return false;
        }
        Expanded e{};
        e.triangle = true;
        e.a = o.points[i0]; e.b = o.points[i1]; e.c = o.points[i2];
        const f3 n = normalize(cross(e.c - e.a, e.b - e.a));                    // Scene.hs:78
        const bool sm = o.hasSmooth ? o.smooth[k] != 0 : false;
        if (o.hasNormals && sm) {
          const int64_t nn = (int64_t)o.normals.size();
          if (i0 >= nn || i1 >= nn || i2 >= nn) { err = "TriangleMesh: normal index out of range"; return false; }
          e.n0 = o.normals[i0]; e.n1 = o.normals[i1]; e.n2 = o.normals[i2];
        } else {
          e.n0 = e.n1 = e.n2 = n;
        }
        e.materialId = &o.materialId;
        out.push_back(e);
      }
      return true;
    }
    case Obj::Subsampled: {
      // ofHundred = ceiling (subsampleSize * 100); pick: take ofHundred, drop the next 100 - ofHundred
      std::vector<Expanded> all;
      for (const Obj& c : o.children) if (!expandObject(c, all, err)) return false;
      const int64_t ofHundred = (int64_t)std::ceil(o.subsampleSize * 100.0f);
      size_t pos = 0;
      while (pos < all.size()) {                                               // pick [] = []
        const size_t take = ofHundred > 0 ? (size_t)ofHundred : 0;             // splitAt n | n <= 0 = ([], xs)
        const size_t end = pos + take < all.size() ? pos + take : all.size();
        for (size_t k = pos; k < end; ++k) out.push_back(all[k]);
        const int64_t drop = 100 - ofHundred;
        pos = end + (drop > 0 ? (size_t)drop : 0);
        if (take == 0 && drop <= 0) break;     // the reference would loop forever here; stop instead
      }
      return true;
    }
    case Obj::Triangle: {
      Expanded e{};
      e.triangle = true;
      e.a = o.p0; e.b = o.p1; e.c = o.p2; e.n0 = o.n0; e.n1 = o.n1; e.n2 = o.n2;
      e.materialId = &o.materialId;
      out.push_back(e);
      return true;
    }
    default: {
      Expanded e{};
      e.triangle = false;
      e.a = o.position; e.radius = o.radius; e.materialId = &o.materialId;
      out.push_back(e);
      return true;
    }
  }
}

void put3(std::vector<float>& v, f3 p) { v.push_back(p.x); v.push_back(p.y); v.push_back(p.z); }

// ---- PNG ------------------------------------------------------------------------------------------
uint32_t crc32(const uint8_t* p, size_t n, uint32_t crc = 0) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

void be32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

bool writeChunk(FILE* f, const char type[4], const std::vector<uint8_t>& data) {
  std::vector<uint8_t> hdr;
  be32(hdr, (uint32_t)data.size());
  std::vector<uint8_t> body(type, type + 4);
  body.insert(body.end(), data.begin(), data.end());
  std::vector<uint8_t> tail;
  be32(tail, crc32(body.data(), body.size()));
  return fwrite(hdr.data(), 1, 4, f) == 4 && fwrite(body.data(), 1, body.size(), f) == body.size() &&
         fwrite(tail.data(), 1, 4, f) == 4;
}

}  // namespace

yahr_scene_desc LoadedScene::desc() const {
  yahr_scene_desc d{};
  d.n_triangles = (uint32_t)triMaterial.size();
  d.tri_p0 = triP0.data(); d.tri_p1 = triP1.data(); d.tri_p2 = triP2.data();
  d.tri_n0 = triN0.data(); d.tri_n1 = triN1.data(); d.tri_n2 = triN2.data();
  d.tri_material = triMaterial.data();
  d.n_spheres = (uint32_t)sphMaterial.size();
  d.sph_center = sphCenter.data(); d.sph_radius = sphRadius.data(); d.sph_material = sphMaterial.data();
  d.prim_order = primOrder.data();
  d.n_materials = (uint32_t)materials.size();
  d.materials = materials7.data();
  d.n_lights = (uint32_t)(lights.size() / 6);
  d.lights = lights.data();
  d.bvh_max_depth = bvhMaxDepth;
  d.split_mode = splitMode;
  return d;
}

int loadYahrr(const std::string& text, LoadedScene& out, std::string& err) {
  out = LoadedScene();
  Parser p;
  if (!lex(text, p.t, err)) { err = "no parse: " + err; return YAHR_ERR_PARSE; }
  std::vector<Obj> objs;
  if (!p.scene(out, objs) || p.t[p.i].kind != kEnd) {
    if (p.err.empty()) p.failAt("trailing input");
    err = "no parse: " + p.err;
    return YAHR_ERR_PARSE;
  }
  // mats = fromList [(id m, shader m) | m <- materials]: a later duplicate id replaces an earlier one
  std::map<std::string, uint32_t> matIndex;
  for (uint32_t m = 0; m < out.materials.size(); ++m) {
    matIndex[out.materials[m].id] = m;
    const SceneMaterial& sm = out.materials[m];
    const float v[7] = {sm.diffuse[0], sm.diffuse[1], sm.diffuse[2], sm.specular[0], sm.specular[1], sm.specular[2],
                        sm.shininess};
    out.materials7.insert(out.materials7.end(), v, v + 7);
  }
  std::vector<Expanded> ex;
  for (const Obj& o : objs)                                       // S.objects s >>= S.expand   (main.hs:44)
    if (!expandObject(o, ex, err)) return YAHR_ERR_PARSE;
  for (const Expanded& e : ex) {
    auto it = matIndex.find(*e.materialId);
    if (it == matIndex.end()) {                                   // Map.! : given key is not an element in the map
      err = "unknown material id \"" + *e.materialId + "\"";
      return YAHR_ERR_UNKNOWN_MATERIAL;
    }
    if (e.triangle) {
      out.primOrder.push_back(0x80000000u | (uint32_t)out.triMaterial.size());
      put3(out.triP0, e.a); put3(out.triP1, e.b); put3(out.triP2, e.c);
      put3(out.triN0, e.n0); put3(out.triN1, e.n1); put3(out.triN2, e.n2);
      out.triMaterial.push_back(it->second);
    } else {
      out.primOrder.push_back((uint32_t)out.sphMaterial.size());
      put3(out.sphCenter, e.a);
      out.sphRadius.push_back(e.radius);
      out.sphMaterial.push_back(it->second);
    }
  }
  return YAHR_OK;
}

int writePngRgb8(const std::string& path, const uint8_t* rgb, int width, int height, std::string& err) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { err = "cannot open " + path + " for writing"; return YAHR_ERR_IO; }
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  bool ok = fwrite(sig, 1, 8, f) == 8;
  std::vector<uint8_t> ihdr;
  be32(ihdr, (uint32_t)width); be32(ihdr, (uint32_t)height);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit RGB
  ok = ok && writeChunk(f, "IHDR", ihdr);
  // raw scanlines (filter byte 0), wrapped in a zlib stream of stored deflate blocks
  const size_t row = (size_t)width * 3 + 1, rawSize = row * (size_t)height;
  std::vector<uint8_t> raw(rawSize);
  for (int y = 0; y < height; ++y) {
    raw[(size_t)y * row] = 0;
    std::memcpy(&raw[(size_t)y * row + 1], rgb + (size_t)y * width * 3, (size_t)width * 3);
  }
  std::vector<uint8_t> z;
  z.reserve(rawSize + rawSize / 65535 * 5 + 16);
  z.push_back(0x78); z.push_back(0x01);
  uint32_t a = 1, b = 0;                                          // adler32
  size_t pos = 0;
  do {
    const size_t n = rawSize - pos < 65535 ? rawSize - pos : 65535;
    z.push_back(pos + n == rawSize ? 1 : 0);
    z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8));
    z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
    for (size_t k = 0; k < n; ++k) { a = (a + raw[pos + k]) % 65521u; b = (b + a) % 65521u; }
    pos += n;
  } while (pos < rawSize);
  be32(z, (b << 16) | a);
  ok = ok && writeChunk(f, "IDAT", z) && writeChunk(f, "IEND", {});
  ok = (fclose(f) == 0) && ok;
  if (!ok) { err = "write error on " + path; return YAHR_ERR_IO; }
  return YAHR_OK;
}

}  // namespace yb
