// yahr_cli.cpp -- the reference's command line (main.hs:28-38, 112-145) on top of libyahr_b200.
//
//   yahr <input.yahrr> <output.png> [-p|--parallel-mode MODE] [+RTS ... [-RTS]]
//
// Same positionals, same flag, same one-line status output ("<N> threads, <B> batches, parallel
// <mode>", main.hs:140-141), exit 0 on success.  MODE is accepted as sequential | eval | par (the
// reference's values, passed by the Blender add-on: compat/blender/render_engine.py:44-48) or gpu;
// every mode renders on the GPU -- there is no CPU path.  GHC RTS options are accepted and
// ignored except -N<k>, which only feeds the status line and the reference's batch-count formula.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/yahr_b200.h"

static int die(const char* what) {
  fprintf(stderr, "yahr: %s: %s\n", what, yahr_b200_last_error());
  return 1;
}

int main(int argc, char** argv) {
  std::vector<std::string> pos;
  std::string mode = "sequential";                       // optFlag "sequential" "parallel-mode" (main.hs:38)
  long threads = 1;                                      // getNumCapabilities without -N
  bool rts = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "+RTS") { rts = true; continue; }
    if (a == "-RTS") { rts = false; continue; }
    if (rts) {                                           // the GHC runtime strips these before main
      if (a.rfind("-N", 0) == 0 && a.size() > 2) threads = std::max(1L, atol(a.c_str() + 2));
      continue;
    }
    if (a == "-p" || a == "--parallel-mode") {
      if (i + 1 >= argc) { fprintf(stderr, "yahr: option %s needs a value\n", a.c_str()); return 1; }
      mode = argv[++i];
      continue;
    }
    if (a.rfind("--parallel-mode=", 0) == 0) { mode = a.substr(16); continue; }
    if (a == "-h" || a == "--help") {
      printf("usage: yahr input output [-p|--parallel-mode sequential|eval|par|gpu] [+RTS -N<k> -RTS]\n");
      return 0;
    }
    pos.push_back(a);
  }
  if (pos.size() != 2) {
    fprintf(stderr, "usage: yahr input output [-p|--parallel-mode MODE]\n");
    return 1;
  }
  if (mode != "sequential" && mode != "eval" && mode != "par" && mode != "gpu") {
    // the reference: non-exhaustive patterns in case (main.hs:134-137)
    fprintf(stderr, "yahr: unknown parallel mode \"%s\"\n", mode.c_str());
    return 1;
  }

  std::ifstream in(pos[0], std::ios::binary);
  if (!in) { fprintf(stderr, "yahr: %s: openFile: does not exist\n", pos[0].c_str()); return 1; }
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string text = ss.str();

  yahr_loaded_scene* loaded = nullptr;
  if (yahr_b200_yahrr_load(text.data(), text.size(), &loaded)) return die("Prelude.read");
  yahr_scene_desc desc;
  yahr_camera cam;
  int depth = 1;
  yahr_b200_yahrr_describe(loaded, &desc, &cam, &depth);
  const int width = (int)std::floor(cam.imW), height = (int)std::floor(cam.imH);   // main.hs:122-123

  yahr_scene* scene = nullptr;
  if (yahr_b200_scene_create(&desc, &scene)) { yahr_b200_yahrr_free(loaded); return die("scene"); }
  yahr_b200_yahrr_free(loaded);

  unsigned char* rgb8 = nullptr;
  const size_t bytes = (size_t)width * height * 3;
  if (width < 1 || height < 1 || yahr_b200_host_alloc(bytes, (void**)&rgb8)) {
    yahr_b200_scene_destroy(scene);
    return die("image");
  }
  yahr_stats st;
  if (yahr_b200_render_rgb8(scene, &cam, depth, 1, 0, rgb8, &st)) {
    yahr_b200_host_free(rgb8);
    yahr_b200_scene_destroy(scene);
    return die("render");
  }
  const long long batches = yahr_b200_num_batches(threads, width, height);
  printf("%ld threads, %lld batches, parallel %s\n", threads, batches, mode.c_str());   // main.hs:140-141
  if (getenv("YAHR_B200_VERBOSE"))
    fprintf(stderr, "yahr_b200: %llu primary + %llu shadow + %llu reflection rays, %.3f ms on the GPU, %.3f ms call\n",
            (unsigned long long)st.n_primary, (unsigned long long)st.n_shadow, (unsigned long long)st.n_secondary,
            st.gpu_ms, st.wall_ms);
  int rc = yahr_b200_write_png_rgb8(pos[1].c_str(), rgb8, width, height);           // savePngImage (main.hs:142)
  yahr_b200_host_free(rgb8);
  yahr_b200_scene_destroy(scene);
  if (rc) return die("savePngImage");
  return 0;
}
