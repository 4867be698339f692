// capi.cu -- the C ABI of libyahr_b200.so (include/yahr_b200.h).  No CPU fallback: every compute
// entry point needs a CUDA device and fails with YAHR_ERR_NO_DEVICE / YAHR_ERR_CUDA otherwise.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/yahr_b200.h"
#include "bvh_build.hpp"
#include "bvh_build_gpu.hpp"
#include "device_types.cuh"
#include "host_scene.hpp"
#include "kernels.hpp"
#include "wide_bvh.hpp"
#include "yahrr_parser.hpp"

using namespace yb;

namespace {

thread_local std::string g_lastError;

int fail(int code, const std::string& msg) {
  g_lastError = msg;
  return code;
}

struct CudaFailure { cudaError_t e; const char* what; const char* file; int line; };

#define CU(call)                                                        \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) throw CudaFailure{e__, #call, __FILE__, __LINE__}; \
  } while (0)

int cudaFail(const CudaFailure& f) {
  int code = f.e == cudaErrorMemoryAllocation ? YAHR_ERR_OUT_OF_MEMORY
             : (f.e == cudaErrorNoDevice || f.e == cudaErrorInsufficientDriver) ? YAHR_ERR_NO_DEVICE
                                                                                : YAHR_ERR_CUDA;
  return fail(code, std::string("CUDA error: ") + cudaGetErrorString(f.e) + " in " + f.what + " (" + f.file + ":" +
                        std::to_string(f.line) + ")");
}

double nowMs() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class T>
T* devUpload(const std::vector<T>& v, uint64_t& bytes) {
  if (v.empty()) return nullptr;
  T* p = nullptr;
  CU(cudaMalloc(&p, v.size() * sizeof(T)));
  CU(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  bytes += v.size() * sizeof(T);
  return p;
}

struct TileSet {
  int4* d_tiles = nullptr;
  uint32_t* d_tileStart = nullptr;
  uint32_t* d_itemPixels = nullptr;   // item -> (u | v << 16), built on the device when the set is created
  uint32_t n = 0, nItems = 0;
  std::vector<int4> hostTiles;        // windows, batch order (row-major over the tile grid)
  std::vector<uint32_t> hostStart;    // n + 1 prefix sums of the tile pixel counts
  uint32_t gridNx = 0;                // tiles per tile-row when the set is the whole image (stride 1)
  // sets made of whole tile ROWS (stride / offset count rows of the tile grid): first tile of every owned row
  // (+ end) and the pixel rows [y0, y1) it covers
  std::vector<uint32_t> rowFirst;
  std::vector<int2> rowY;
  // streamed host output: image row -> owned tile row, pixels per tile row
  unsigned short* d_rowOfV = nullptr;
  uint32_t* d_rowItems = nullptr;
};

}  // namespace

// Host-buffer entry, experiments only (YAHR_B200_HOST_MEASURE=1; the default is the static rule in renderHost): measured
// choice between the copy-engine bands and the streamed rows.  Calls 0-2 of a key use the bands, calls 3-5 the streamed
// rows; the first call of each kind is a warm-up, the faster of the other two counts.  The streamed rows are kept only
// when they win by more than 3 %.
struct HostStrategy {
  int calls = 0;
  double msBands = 1e30, msStream = 1e30;
  bool exploring() const { return calls < 6; }
  bool wantStream() const { return calls < 3 ? false : (calls < 6 ? true : msStream < 0.97 * msBands); }
  void record(bool streamed, double ms) {
    if (!exploring()) return;
    if (!streamed && (calls == 1 || calls == 2) && ms < msBands) msBands = ms;
    if (streamed && (calls == 4 || calls == 5) && ms < msStream) msStream = ms;
    ++calls;
  }
};

struct yahr_scene {
  int device = 0;
  DeviceScene dev{};
  float4* d_nodes = nullptr;
  float4* d_wide = nullptr;              // 4-wide collapse of d_nodes (NULL: the root is a leaf, or not built)
  float4 *d_cwide = nullptr, *d_leafBox = nullptr, *d_multiBox = nullptr;   // compressed copy of d_wide + exact leaf boxes
  int preferCompressed = 0;              // per-scene choice: walk the compressed nodes by default
  float4* d_prims = nullptr;
  float4* d_normals = nullptr;
  uint2* d_multi = nullptr;
  float4* d_materials = nullptr;
  float4* d_lights = nullptr;
  float4* d_areaLights = nullptr;
  uint32_t nTriangles = 0;                // spheres only: the (short) leaf code runs at a lower parking threshold
  unsigned long long* d_counters = nullptr;
  uint32_t* d_order = nullptr;            // primitive ID per DFS position (inspection)
  yahr_scene_info info{};
  // per (width, height, stride, offset) tile lists, uploaded once
  std::map<std::tuple<int, int, int, int, int>, TileSet> tiles;
  // wavefront scratch (grown on demand): hit records, shadow queue, work counters, spp buffers
  // two slots so that consecutive bands of the host-buffer entry can be in flight on two streams
  float4 *wfQ0[2] = {nullptr, nullptr}, *wfQ1[2] = {nullptr, nullptr}, *wfQ2[2] = {nullptr, nullptr};
  unsigned char* wfVis[2] = {nullptr, nullptr};
  uint2* wfHits[2] = {nullptr, nullptr};        // three-kernel set: hit records between k_wf_trace and k_wf_shade
  size_t wfHitEntries[2] = {0, 0};
  size_t wfEntries[2] = {0, 0};
  unsigned long long* d_workStats = nullptr;    // counting build: 16 work counters
  int preferBinary = 0;                         // per-scene choice of the walk: 1 = binary tree, 0 = its 4-wide collapse
  uint32_t* wfWork = nullptr;              // 2 x 8 counters
  float *wfSampleBuf = nullptr, *wfAccum = nullptr;
  size_t wfPixels = 0, wfAccumPixels = 0;  // capacities: wfSampleBuf holds wfPixels per-sample pixels, wfAccum wfAccumPixels
  int numSMs = 148;
  // frame buffers of the host-buffer entry (grown on demand)
  float* d_rgb = nullptr;
  uint32_t* d_primid = nullptr;
  size_t framePixels = 0;
  unsigned char* d_rgb8 = nullptr;
  size_t frame8Pixels = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evIn = nullptr;
  cudaEvent_t phaseEv[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t renderStream[2] = {nullptr, nullptr}, copyStream = nullptr;   // host-buffer entry: render / D2H overlap
  std::vector<cudaEvent_t> bandEvents;
  // host-buffer entry: shadow probes per band of the previous call with the same band layout (a cost estimate that
  // orders the bands of the next call: cheap ones first)
  uint32_t* d_bandProbes = nullptr;
  std::vector<uint32_t> bandProbes;
  std::tuple<int, int, int, int, int> bandKey{0, 0, 0, 0, 0};
  float lastHostGpuMs = 0.0f;
  // host-buffer entry, streamed rows: per-row counters, completion flags in mapped host memory, frame sequence number
  uint32_t* d_rowDone = nullptr;
  uint32_t* h_rowFlags = nullptr;
  uint32_t* d_rowFlags = nullptr;
  uint32_t rowCap = 0, rowSeq = 0;
  std::map<std::tuple<int, int, int, int, int>, HostStrategy> hostStrategy;

  ~yahr_scene() {
    cudaFree(d_cwide); cudaFree(d_leafBox); cudaFree(d_multiBox);
    cudaFree(d_nodes); cudaFree(d_wide); cudaFree(d_prims); cudaFree(d_normals); cudaFree(d_multi); cudaFree(d_materials);
    cudaFree(d_lights); cudaFree(d_areaLights); cudaFree(d_counters); cudaFree(d_order); cudaFree(d_rgb); cudaFree(d_primid); cudaFree(d_rgb8);
    for (auto& kv : tiles) {
      cudaFree(kv.second.d_tiles); cudaFree(kv.second.d_tileStart); cudaFree(kv.second.d_itemPixels);
      cudaFree(kv.second.d_rowOfV); cudaFree(kv.second.d_rowItems);
    }
    cudaFree(d_rowDone);
    if (h_rowFlags) cudaFreeHost(h_rowFlags);
    for (int k = 0; k < 2; ++k) { cudaFree(wfQ0[k]); cudaFree(wfQ1[k]); cudaFree(wfQ2[k]); cudaFree(wfVis[k]); cudaFree(wfHits[k]); }
    cudaFree(d_workStats);
    cudaFree(wfWork); cudaFree(wfSampleBuf); cudaFree(wfAccum); cudaFree(d_bandProbes);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (evIn) cudaEventDestroy(evIn);
    for (auto e : phaseEv) if (e) cudaEventDestroy(e);
    for (auto e : bandEvents) cudaEventDestroy(e);
    for (auto st : renderStream) if (st) cudaStreamDestroy(st);
    if (copyStream) cudaStreamDestroy(copyStream);
  }
};

struct yahr_host_bvh {
  HostBvh bvh;
};

struct yahr_loaded_scene {
  LoadedScene scene;
};

namespace {

const TileSet& tilesFor(yahr_scene* sc, int w, int h, int stride, int offset, int byRows = 0) {
  auto key = std::make_tuple(w, h, stride, offset, byRows);
  auto it = sc->tiles.find(key);
  if (it != sc->tiles.end()) return it->second;
  // samplePoints = squareBatches width height nBatches (main.hs:128-131).  numThreads only enters
  // through max (32 * numThreads) width; the GPU path uses 1 (tiles are a scheduling unit, the
  // image does not depend on them).
  const int64_t nBatches = numBatches(1, w, h);
  std::vector<int4> host;
  std::vector<uint32_t> start(1, 0u);
  TileSet ts;
  int64_t gnx = nBatches, gny = 1;                     // (nx, ny) = loop count 1   (Sampling.hs:11-15)
  while (gnx % 2 == 0 && 2 * (int64_t)w * gny < (int64_t)h * gnx) { gnx /= 2; gny *= 2; }
  if (!byRows) {
    for (int64_t b = offset; b < nBatches; b += stride) {
      TileWindow t = batchWindow(w, h, b, nBatches);
      if (t.x1 > t.x0 && t.y1 > t.y0) {
        host.push_back(make_int4(t.x0, t.y0, t.x1, t.y1));
        start.push_back(start.back() + (uint32_t)((t.x1 - t.x0) * (t.y1 - t.y0)));
      }
    }
  } else {
    // tile `num` sits in row num `quot` nx of the grid (Sampling.hs:16).  byRows = 1: rows offset, offset + stride, ...
    // byRows = 2 (host-buffer shards): the rows are dealt in BLOCKS of up to four consecutive rows (block b -> shard
    // b mod stride), so that a shard's device-to-host copies are few and large (every copy costs ~5 us of copy-engine
    // idle time) while every shard still gets at least eight blocks spread over the image.
    int64_t block = 1;
    if (byRows == 2) { block = gny / (8 * (int64_t)stride); block = block < 1 ? 1 : (block > 4 ? 4 : block); }
    for (int64_t j = 0; j < gny; ++j) {
      if ((j / block) % stride != offset) continue;
      const uint32_t firstOfRow = (uint32_t)host.size();
      int y0 = 0, y1 = 0;
      for (int64_t i = 0; i < gnx; ++i) {
        TileWindow t = batchWindow(w, h, j * gnx + i, nBatches);
        if (t.x1 > t.x0 && t.y1 > t.y0) {
          host.push_back(make_int4(t.x0, t.y0, t.x1, t.y1));
          start.push_back(start.back() + (uint32_t)((t.x1 - t.x0) * (t.y1 - t.y0)));
          y0 = t.y0; y1 = t.y1;
        }
      }
      if (host.size() > firstOfRow) { ts.rowFirst.push_back(firstOfRow); ts.rowY.push_back(make_int2(y0, y1)); }
    }
    ts.rowFirst.push_back((uint32_t)host.size());
  }
  ts.n = (uint32_t)host.size();
  ts.nItems = start.back();
  ts.hostTiles = host;
  ts.hostStart = start;
  if (stride == 1) ts.gridNx = (ts.n == (uint32_t)nBatches) ? (uint32_t)gnx : 0;   // 0 when empty windows were dropped
  uint64_t bytes = 0;
  ts.d_tiles = devUpload(host, bytes);
  ts.d_tileStart = devUpload(start, bytes);
  if (byRows && !ts.rowY.empty() && ts.rowY.size() < 0xFFFFu) {
    std::vector<unsigned short> rowOfV((size_t)h, (unsigned short)0xFFFF);
    std::vector<uint32_t> rowItems(ts.rowY.size());
    for (size_t r = 0; r < ts.rowY.size(); ++r) {
      rowItems[r] = start[ts.rowFirst[r + 1]] - start[ts.rowFirst[r]];
      for (int y = ts.rowY[r].x; y < ts.rowY[r].y; ++y) rowOfV[(size_t)y] = (unsigned short)r;
    }
    ts.d_rowOfV = devUpload(rowOfV, bytes);
    ts.d_rowItems = devUpload(rowItems, bytes);
  }
  if (ts.nItems && w <= 0xFFFF && h <= 0xFFFF && !getenv("YAHR_B200_NO_PIXEL_TABLE")) {
    CU(cudaMalloc(&ts.d_itemPixels, (size_t)ts.nItems * sizeof(uint32_t)));
    WavefrontParams T{};
    T.base.tiles = ts.d_tiles; T.base.nTiles = ts.n; T.base.width = w; T.base.height = h;
    T.tileStart = ts.d_tileStart; T.nItems = ts.nItems; T.itemBase = 0;
    CU(launchPixelTable(T, ts.d_itemPixels, 0));
    CU(cudaStreamSynchronize(0));
  }
  return sc->tiles.emplace(key, ts).first->second;
}

// Defaults of the wavefront set, from the sweeps in profiles/ (r2): see DESIGN.md section 4.
constexpr uint32_t kDefaultPersist = 0u;         // device-resident frames: 1 = k_wf_persist, 0 = k_wf_primary + k_wf_shadow
constexpr uint32_t kDefaultStackShared = 0u;     // traversal-stack entries per lane in shared memory
constexpr uint32_t kDefaultSplit = 0u;           // 1 = three-kernel set (trace / shade / shadow)
constexpr uint32_t kFusedMaxItems = 2500000u;    // one-slot frames up to this many work items per launch use k_wf_fused

// Everything a frame needs, validated once; tiles are then enqueued in one or several ranges.
struct FramePlan {
  CameraSetup cs;
  const TileSet* ts = nullptr;
  RenderParams P{};
  WavefrontParams W{};
  bool wavefront = false;
  uint32_t entriesPerItem = 1;
  uint32_t samplesPerLaunch = 1;      // spp > 1: samples of every pixel traced by one launch
  bool counted = false;               // the counting build of the kernels (yahr_b200_render_device_counted)
};

int planFrame(yahr_scene* sc, const yahr_camera* cam, const yahr_render_opts* opts, float* d_rgb, uint32_t* d_primid,
              FramePlan& plan) {
  CameraSetup& cs = plan.cs;
  std::string err;
  int rc = setupCamera(cam, cs, err);
  if (rc) return fail(rc, err);
  if (opts->spp < 1) return fail(YAHR_ERR_INVALID_ARGUMENT, "spp must be >= 1");
  if (opts->recursion_depth < 0 || opts->recursion_depth > YAHR_B200_MAX_RECURSION)
    return fail(YAHR_ERR_INVALID_ARGUMENT, "recursion_depth out of range [0, 16]");
  if (opts->tile_stride < 1 || opts->tile_offset < 0 || opts->tile_offset >= opts->tile_stride)
    return fail(YAHR_ERR_INVALID_ARGUMENT, "tile_stride/tile_offset invalid");
  if (opts->traversal != YAHR_TRAVERSAL_REFERENCE && opts->traversal != YAHR_TRAVERSAL_ORDERED)
    return fail(YAHR_ERR_INVALID_ARGUMENT, "unknown traversal mode");
  // kernel selection: 0 = default (wavefront set for direct lighting, megakernel for recursion
  // depth != 1), 1 = megakernel, 2 = wavefront
  // recursion depth >= 2 has per-batch wavefront kernels for the reference's own scene shape: point lights only (area
  // lights are the extension), 4-wide tree, reference traversal order (k_wf_fused_depth*); everything else recursive is
  // the megakernel
  const bool depthKernel = opts->recursion_depth >= 2 && sc->dev.nLights >= 1 && sc->dev.nAreaLights == 0 &&
                           sc->dev.nSlots == sc->dev.nLights && sc->dev.wide != nullptr &&
                           opts->traversal == YAHR_TRAVERSAL_REFERENCE &&
                           !getenv("YAHR_B200_NO_DEPTH_KERNEL");
  plan.wavefront = opts->kernel == 2 || (opts->kernel == 0 && (opts->recursion_depth == 1 || depthKernel));
  if (opts->kernel == 2 && opts->recursion_depth != 1 && !depthKernel)
    return fail(YAHR_ERR_INVALID_ARGUMENT,
                "the wavefront kernels handle recursion_depth 1, or >= 2 with point lights only (no area lights)");

  // reserved[1] = 1: tile_stride / tile_offset count whole ROWS of the tile grid (host-buffer shards)
  const TileSet& ts = tilesFor(sc, cs.width, cs.height, opts->tile_stride, opts->tile_offset,
                               (opts->reserved[1] == 1 || opts->reserved[1] == 2) ? opts->reserved[1] : 0);
  plan.ts = &ts;
  RenderParams& P = plan.P;
  P.sc = sc->dev;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) P.vtf[4 * i + j] = cs.vtf.m[i][j];
  for (int i = 0; i < 3; ++i) P.origin[i] = cs.origin[i];
  P.focal = cam->focalLength;
  P.width = cs.width; P.height = cs.height;
  P.depth = opts->recursion_depth; P.spp = opts->spp; P.seed = opts->seed;
  P.traversal = opts->traversal;
  P.tiles = ts.d_tiles; P.nTiles = ts.n;
  P.rgb = d_rgb; P.primid = d_primid;
  P.counters = sc->d_counters;

  if (plan.wavefront) {
    WavefrontParams& W = plan.W;
    const uint32_t nL = sc->dev.nSlots;          // light slots: point lights + every area-light sample
    plan.entriesPerItem = nL > 1 ? nL : 1;
    const size_t px = (size_t)cs.width * cs.height;
    // spp > 1: one launch traces several samples of every pixel (fewer, larger launches: the persistent kernels'
    // ramp-up and tail are paid once per launch).  As many as fit a scratch budget of ~6 GB for the shadow queue.
    plan.samplesPerLaunch = 1;
    if (opts->spp > 1) {
      const double perSample = (double)ts.nItems * plan.entriesPerItem * 49.0 + (double)px * 12.0;
      double fit = perSample > 0 ? 6.0e9 / perSample : 1.0;
      const double idxLimit = 2.0e9 / ((double)(ts.nItems ? ts.nItems : 1) * plan.entriesPerItem);    // 32-bit entry indices
      if (fit > idxLimit) fit = idxLimit;
      uint32_t k = fit < 1.0 ? 1u : (fit > 16.0 ? 16u : (uint32_t)fit);
      if (const char* env = getenv("YAHR_B200_SAMPLES_PER_LAUNCH")) k = (uint32_t)atoi(env) < 1u ? 1u : (uint32_t)atoi(env);
      if (k > (uint32_t)opts->spp) k = (uint32_t)opts->spp;
      plan.samplesPerLaunch = k;
    }
    if ((double)ts.nItems * plan.entriesPerItem * plan.samplesPerLaunch >= 4.0e9 || (double)px * plan.samplesPerLaunch >= 4.0e9)
      return fail(YAHR_ERR_INVALID_ARGUMENT, "resolution x light slots exceeds the 32-bit shadow-queue index space");
    // the two scratch buffers have separate capacities: a larger image traced with fewer samples per launch can fit
    // the per-sample buffer of an earlier call and still outgrow the accumulator
    if (opts->spp > 1 && px * plan.samplesPerLaunch > sc->wfPixels) {
      CU(cudaDeviceSynchronize());
      cudaFree(sc->wfSampleBuf); sc->wfSampleBuf = nullptr; sc->wfPixels = 0;
      CU(cudaMalloc(&sc->wfSampleBuf, px * plan.samplesPerLaunch * 3 * sizeof(float)));
      sc->wfPixels = px * plan.samplesPerLaunch;
    }
    if (opts->spp > 1 && px > sc->wfAccumPixels) {
      CU(cudaDeviceSynchronize());
      cudaFree(sc->wfAccum); sc->wfAccum = nullptr; sc->wfAccumPixels = 0;
      CU(cudaMalloc(&sc->wfAccum, px * 3 * sizeof(float)));
      sc->wfAccumPixels = px;
    }
    if (!sc->wfWork) {
      CU(cudaMalloc(&sc->wfWork, 16 * sizeof(uint32_t)));
      CU(cudaMemset(sc->wfWork, 0, 16 * sizeof(uint32_t)));
    }
    W.base = P;
    // light slots: 1 -> compacted queue, every pixel stored once; 2 -> compacted queue + float atomics (order-free for
    // two addends; YAHR_B200_TWO_SLOT=0 keeps the dense path); 3 or more -> dense entries + in-order resolve
    // The atomics land in the per-sample frames of the launch; once those are far larger than L2 they cost more than the
    // dense path's streaming resolve (measured, profiles/r2f, r2aa: C2 + area light, 1920x1080, 16 samples per launch =
    // 0.40 GB: 14.7 -> 12.9 ms; C5 + area light, 3840x2160, 6 per launch = 0.60 GB: 75 -> 77 ms at any number of samples
    // per launch), hence the size limit.
    static const int envTwoSlot = getenv("YAHR_B200_TWO_SLOT") ? atoi(getenv("YAHR_B200_TWO_SLOT")) : -1;
    const double sampleFrameBytes = (double)px * plan.samplesPerLaunch * 12.0;
    W.twoSlot = (nL == 2 && (envTwoSlot >= 0 ? envTwoSlot != 0 : sampleFrameBytes <= 0.45e9)) ? 1u : 0u;
    W.tileStart = ts.d_tileStart; W.nItems = ts.nItems; W.itemBase = 0; W.sample = 0; W.dense = (nL > 1 && !W.twoSlot) ? 1u : 0u;
    W.itemPixels = ts.d_itemPixels;
    // tuning knobs (opts->reserved[0]): bits 0-7 leaf-parking threshold (0 = default), bits 16-23 CTAs/SM
    static const uint32_t envTune = getenv("YAHR_B200_TUNE") ? (uint32_t)strtoul(getenv("YAHR_B200_TUNE"), nullptr, 0) : 0u;
    const uint32_t tune = opts->reserved[0] ? (uint32_t)opts->reserved[0] : envTune;
    // leaf parking: the triangle test is long (wait for 12 lanes), the sphere test short (6); profiles/r1k
    W.leafThreshold = (tune & 0xFF) ? (tune & 0xFF) : (sc->nTriangles ? 12u : 6u);
    W.blocksPerSM = (tune >> 16) & 0xFF;
    W.capRegisters = ((tune >> 8) & 1u) ^ 1u;      // default: capped (bit 8 set = uncapped)
    W.packed = ((tune >> 9) & 1u) ^ 1u;            // default: packed node step (bit 9 set = generic)
    // the walk: bit 10 set = binary tree, bit 14 set = its 4-wide collapse, neither = the scene's own choice (attachWideTree)
    W.wideTree = (tune & 0x400u) ? 0u : ((tune & 0x4000u) ? 1u : (sc->preferBinary ? 0u : 1u));
    W.leafRun = ((tune >> 11) & 1u) ^ 1u;          // default: on (bit 11 set = one leaf per leaf phase)
    // bit 30: the compressed 64-byte wide nodes, bit 31: the exact 128-byte ones, neither: the scene's own choice
    W.compressed = (tune & 0x80000000u) ? 0u : ((tune & 0x40000000u) ? 1u : (sc->preferCompressed ? 1u : 0u));
    // bit 12 set = the per-batch kernel k_wf_fused (one light slot), 13: its 72-register build; bit 29 = the two-kernel set.
    // Neither: by size.  A frame of up to kFusedMaxItems work items is dominated by the ramp-up and the tail of the
    // persistent kernels, and one launch beats two (profiles/r2p: C1 0.059 -> 0.048 ms, C2 0.532 -> 0.463 ms, a 1/8 share of
    // C4 0.354 -> 0.325 ms); above it the two-kernel set's compacted any-hit walks win (C3 1.01 vs 1.27 ms, soup 10.0 vs 11.1).
    W.fused = (tune >> 12) & 3u;
    if (!W.fused && !(tune & 0x20000000u) && opts->recursion_depth == 1 &&
        (double)ts.nItems * plan.samplesPerLaunch <= (double)kFusedMaxItems)
      W.fused = 1u;
    // bit 15: one light slot -> ONE persistent kernel with committed probe chunks (k_wf_persist); bit 29 = the two-kernel set
    static const int envPersist = getenv("YAHR_B200_PERSIST") ? atoi(getenv("YAHR_B200_PERSIST")) : -1;
    W.persist = (tune & 0x20000000u) ? 0u : ((tune & 0x8000u) ? 1u : (envPersist >= 0 ? (uint32_t)envPersist : kDefaultPersist));
    // bit 28: the three-kernel set (k_wf_trace -> k_wf_shade -> k_wf_shadow) instead of k_wf_primary -> k_wf_shadow
    static const int envSplit = getenv("YAHR_B200_SPLIT") ? atoi(getenv("YAHR_B200_SPLIT")) : -1;
    W.split = (tune & 0x10000000u) ? 1u : (envSplit >= 0 ? (uint32_t)envSplit : kDefaultSplit);
    // bits 24-27: traversal-stack entries per lane in shared memory: 0 = default, 1 = none (all local), 8, 12
    static const int envSh = getenv("YAHR_B200_STACK_SH") ? atoi(getenv("YAHR_B200_STACK_SH")) : -1;
    const uint32_t shBits = (tune >> 24) & 0xFu;
    W.stackShared = shBits ? (shBits == 1u ? 0u : shBits) : (envSh >= 0 ? (uint32_t)envSh : kDefaultStackShared);
    W.sampleOut = d_rgb; W.sampleBuf = sc->wfSampleBuf; W.accum = sc->wfAccum;
    W.samplesPerLaunch = plan.samplesPerLaunch;
  }
  return YAHR_OK;
}

// Enqueues the kernels for tiles [first, first + count) of the plan's tile set.
void enqueueTiles(yahr_scene* sc, const FramePlan& plan, uint32_t first, uint32_t count, cudaStream_t stream,
                  uint32_t* launches, cudaEvent_t* phaseEv, int slot = 0, uint32_t* bandStat = nullptr) {
  if (count == 0) return;
  const TileSet& ts = *plan.ts;
  if (plan.wavefront) {
    WavefrontParams W = plan.W;
    // (the recursion kernels keep their probes in registers: no queue)
    size_t entries = plan.P.depth != 1 ? 0 : (size_t)(ts.hostStart[first + count] - ts.hostStart[first]) *
                                                 plan.entriesPerItem * plan.samplesPerLaunch;
    // k_wf_persist keeps a 256-entry ring per CTA in q0 / q2 instead (at most 16 CTAs per SM)
    if (plan.P.depth == 1 && W.persist && entries < (size_t)sc->numSMs * 16 * 256) entries = (size_t)sc->numSMs * 16 * 256;
    if (entries > sc->wfEntries[slot]) {          // grows only on the first frame of a given size
      CU(cudaDeviceSynchronize());
      cudaFree(sc->wfQ0[slot]); cudaFree(sc->wfQ1[slot]); cudaFree(sc->wfQ2[slot]); cudaFree(sc->wfVis[slot]);
      sc->wfQ0[slot] = sc->wfQ1[slot] = sc->wfQ2[slot] = nullptr; sc->wfVis[slot] = nullptr; sc->wfEntries[slot] = 0;
      CU(cudaMalloc(&sc->wfQ0[slot], entries * sizeof(float4)));
      CU(cudaMalloc(&sc->wfQ1[slot], entries * sizeof(float4)));
      CU(cudaMalloc(&sc->wfQ2[slot], entries * sizeof(float4)));
      CU(cudaMalloc(&sc->wfVis[slot], entries));
      sc->wfEntries[slot] = entries;
    }
    W.hits = nullptr;
    if (W.split && plan.P.depth == 1) {
      const size_t nItemsL = ts.hostStart[first + count] - ts.hostStart[first];
      const size_t hitEntries = ((nItemsL + 31) & ~(size_t)31) * plan.samplesPerLaunch;
      if (hitEntries > sc->wfHitEntries[slot]) {
        CU(cudaDeviceSynchronize());
        cudaFree(sc->wfHits[slot]); sc->wfHits[slot] = nullptr; sc->wfHitEntries[slot] = 0;
        CU(cudaMalloc(&sc->wfHits[slot], hitEntries * sizeof(uint2)));
        sc->wfHitEntries[slot] = hitEntries;
      }
      W.hits = sc->wfHits[slot];
    }
    W.q0 = sc->wfQ0[slot]; W.q1 = sc->wfQ1[slot]; W.q2 = sc->wfQ2[slot];
    W.visibility = W.dense ? sc->wfVis[slot] : nullptr;
    W.work = sc->wfWork + 8 * slot;
    W.bandStat = bandStat;
    W.base.tiles = ts.d_tiles + first; W.base.nTiles = count;
    W.tileStart = ts.d_tileStart + first;
    W.itemBase = ts.hostStart[first];
    W.nItems = ts.hostStart[first + count] - ts.hostStart[first];
    if (plan.counted) CU(counted::launchWavefrontCounted(W, sc->numSMs, stream, launches, phaseEv));
    else CU(launchWavefront(W, sc->numSMs, stream, launches, phaseEv));
  } else {
    RenderParams P = plan.P;
    P.tiles = ts.d_tiles + first; P.nTiles = count;
    CU(launchRenderMega(P, stream, launches));
  }
}

int renderCommon(yahr_scene* sc, const yahr_camera* cam, const yahr_render_opts* opts, float* d_rgb,
                 uint32_t* d_primid, cudaStream_t stream, yahr_stats* stats, bool sync, uint64_t* workCounts = nullptr) {
  const double w0 = nowMs();
  FramePlan plan;
  int rc = planFrame(sc, cam, opts, d_rgb, d_primid, plan);
  if (rc) return rc;
  if (workCounts) {
    if (!plan.wavefront || opts->recursion_depth != 1)
      return fail(YAHR_ERR_INVALID_ARGUMENT, "the counting build covers the wavefront set at recursion_depth 1");
    if (!sc->d_workStats) CU(cudaMalloc(&sc->d_workStats, 16 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(sc->d_workStats, 0, 16 * sizeof(unsigned long long), stream));
    plan.counted = true;
    plan.W.workStats = sc->d_workStats;
  }
  const TileSet& ts = *plan.ts;
  uint32_t launches = 0;
  CU(cudaMemsetAsync(sc->d_counters, 0, 3 * sizeof(unsigned long long), stream));
  if (stats) CU(cudaEventRecord(sc->ev0, stream));
  enqueueTiles(sc, plan, 0, ts.n, stream, &launches, stats ? sc->phaseEv : nullptr);
  if (stats) CU(cudaEventRecord(sc->ev1, stream));
  if (stats) {
    unsigned long long c[3];
    CU(cudaMemcpyAsync(c, sc->d_counters, sizeof(c), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    std::memset(stats, 0, sizeof(*stats));
    stats->n_primary = c[0]; stats->n_shadow = c[1]; stats->n_secondary = c[2];
    stats->gpu_ms = ms;
    if (plan.wavefront && ts.n) {
      for (int k = 0; k < 3; ++k) {
        float pm = 0;
        CU(cudaEventElapsedTime(&pm, sc->phaseEv[k], sc->phaseEv[k + 1]));
        stats->phase_ms[k] = pm;
      }
    } else {
      stats->phase_ms[0] = ms;
    }
    stats->launches = launches;
    stats->tiles = ts.n;
    stats->wall_ms = nowMs() - w0;
  } else if (sync) {
    CU(cudaStreamSynchronize(stream));
  }
  if (workCounts) {
    CU(cudaStreamSynchronize(stream));
    CU(cudaMemcpy(workCounts, sc->d_workStats, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  return YAHR_OK;
}

// Collapses the device-resident binary tree into 4-wide nodes (wide_bvh.cu); shared by both build paths.
// A tree whose wide walk could outgrow the traversal stack keeps the binary walk only.
void attachWideTree(yahr_scene* sc) {
  if (getenv("YAHR_B200_NO_WIDE") || !sc->d_nodes || sc->info.n_nodes == 0) return;
  const double t0 = nowMs();
  WideBuildOutput wo;
  if (!buildWideOnDevice(sc->d_nodes, sc->info.n_nodes, sc->info.depth, wo)) throw CudaFailure{wo.error, wo.where, __FILE__, __LINE__};
  if (wo.wide && wo.stackNeed < (uint32_t)YAHR_B200_MAX_STACK) {   // one scratch slot past the top (wideStep)
    sc->d_wide = wo.wide;
    sc->dev.wide = wo.wide;
    sc->info.n_wide_nodes = wo.nWide;
    sc->info.device_bytes += (uint64_t)wo.nWide * kWideNodeVec * sizeof(float4);
    if (!getenv("YAHR_B200_NO_COMPRESSED")) {
      CompressedWideOutput co;
      if (!compressWideOnDevice(wo.wide, wo.nWide, sc->info.n_primitives, sc->info.n_multi_leaves, co))
        throw CudaFailure{co.error, co.where, __FILE__, __LINE__};
      if (co.nodes) {
        sc->d_cwide = co.nodes; sc->d_leafBox = co.leafBox; sc->d_multiBox = co.multiBox;
        sc->dev.cwide = co.nodes; sc->dev.leafBox = co.leafBox; sc->dev.multiBox = co.multiBox;
        sc->info.device_bytes += co.bytes;
      }
    }
  } else {
    cudaFree(wo.wide);
  }
  sc->info.build_ms += nowMs() - t0;
}

// Extension: quad area lights (include/yahr_b200.h).  normal = norm (edge1 x edge2) and flux = (area / samples) @*
// radiance are computed here once, in binary32, cross -> sqrt (dot n n) -> scale, in that order.
int uploadAreaLights(yahr_scene* sc, const yahr_scene_desc* desc, uint64_t& bytes) {
  sc->dev.areaLights = nullptr; sc->dev.nAreaLights = 0; sc->dev.nSlots = desc->n_lights;
  if (desc->n_area_lights == 0) return YAHR_OK;
  if (!desc->area_lights) return fail(YAHR_ERR_INVALID_ARGUMENT, "area_lights missing");
  std::vector<float4> recs(5 * (size_t)desc->n_area_lights);
  uint64_t slots = desc->n_lights;
  for (uint32_t a = 0; a < desc->n_area_lights; ++a) {
    const float* f = desc->area_lights + 13 * (size_t)a;
    for (int k = 0; k < 13; ++k)
      if (!std::isfinite(f[k])) return fail(YAHR_ERR_NON_FINITE_INPUT, "non-finite area light");
    if (!(f[12] >= 1.0f && f[12] <= 4096.0f)) return fail(YAHR_ERR_INVALID_ARGUMENT, "area light samples must be in [1, 4096]");
    const uint32_t samples = (uint32_t)f[12];
    const f3 corner{f[0], f[1], f[2]}, e1{f[3], f[4], f[5]}, e2{f[6], f[7], f[8]}, radiance{f[9], f[10], f[11]};
    const f3 n = cross(e1, e2);
    const float area = std::sqrt(dot(n, n));
    if (!(area > 0.0f)) return fail(YAHR_ERR_INVALID_ARGUMENT, "degenerate area light");
    const f3 normal = (1.0f / area) * n;
    const f3 flux = (area / (float)samples) * radiance;
    float sbits;
    std::memcpy(&sbits, &samples, 4);
    recs[5 * a + 0] = make_float4(corner.x, corner.y, corner.z, sbits);
    recs[5 * a + 1] = make_float4(e1.x, e1.y, e1.z, 0.0f);
    recs[5 * a + 2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
    recs[5 * a + 3] = make_float4(normal.x, normal.y, normal.z, 0.0f);
    recs[5 * a + 4] = make_float4(flux.x, flux.y, flux.z, 0.0f);
    slots += samples;
  }
  if (slots > 65535) return fail(YAHR_ERR_INVALID_ARGUMENT, "more than 65535 light slots");
  sc->d_areaLights = devUpload(recs, bytes);
  sc->dev.areaLights = sc->d_areaLights; sc->dev.nAreaLights = desc->n_area_lights; sc->dev.nSlots = (uint32_t)slots;
  return YAHR_OK;
}

// Small per-scene tables (materials, lights), events, counters: shared by both build paths.
void uploadSmallTables(yahr_scene* sc, const yahr_scene_desc* desc, uint64_t& bytes) {
  std::vector<float4> mats(2 * (size_t)desc->n_materials), lights(2 * (size_t)desc->n_lights);
  for (uint32_t m = 0; m < desc->n_materials; ++m) {
    const float* s = desc->materials + 7 * (size_t)m;
    mats[2 * m + 0] = make_float4(s[0], s[1], s[2], s[6]);
    mats[2 * m + 1] = make_float4(s[3], s[4], s[5], 0.0f);
  }
  for (uint32_t l = 0; l < desc->n_lights; ++l) {
    const float* s = desc->lights + 6 * (size_t)l;
    lights[2 * l + 0] = make_float4(s[0], s[1], s[2], 0.0f);
    lights[2 * l + 1] = make_float4(s[3], s[4], s[5], 0.0f);
  }
  sc->d_materials = devUpload(mats, bytes);
  sc->d_lights = devUpload(lights, bytes);
  CU(cudaMalloc(&sc->d_counters, 8 * sizeof(unsigned long long)));
  CU(cudaEventCreate(&sc->ev0));
  CU(cudaEventCreate(&sc->ev1));
  for (auto& e : sc->phaseEv) CU(cudaEventCreate(&e));
  sc->dev.materials = sc->d_materials; sc->dev.lights = sc->d_lights; sc->dev.nLights = desc->n_lights;
}

// Copies one caller array into the input arena (256-byte aligned slices of one stream-ordered allocation).
template <class T>
const T* arenaCopy(const T* host, size_t count, char* base, size_t& off) {
  if (!host || count == 0) return nullptr;
  off = (off + 255) & ~(size_t)255;
  T* p = reinterpret_cast<T*>(base + off);
  off += count * sizeof(T);
  CU(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, 0));
  return p;
}

// Device build path: upload the caller's arrays as they are, build on the GPU (bvh_build_gpu.cu).
// Returns YAHR_OK, a positive error code, or -1 when the device builder declines (host builder next).
int createSceneOnDevice(const yahr_scene_desc* d, yahr_scene** out) {
  const uint64_t n = (uint64_t)d->n_spheres + d->n_triangles;
  if (n >= 0x40000000ull) return fail(YAHR_ERR_INVALID_ARGUMENT, "more than 2^30 primitives");
  if (d->n_triangles && !(d->tri_p0 && d->tri_p1 && d->tri_p2 && d->tri_n0 && d->tri_n1 && d->tri_n2))
    return fail(YAHR_ERR_INVALID_ARGUMENT, "triangle arrays missing");
  if (d->n_spheres && !(d->sph_center && d->sph_radius)) return fail(YAHR_ERR_INVALID_ARGUMENT, "sphere arrays missing");
  if (d->n_materials && !d->materials) return fail(YAHR_ERR_INVALID_ARGUMENT, "materials missing");
  if (d->n_lights && !d->lights) return fail(YAHR_ERR_INVALID_ARGUMENT, "lights missing");
  char* inputArena = nullptr;
  yahr_scene* sc = nullptr;
  GpuBuildOutput bo;
  auto cleanup = [&]() { if (inputArena) cudaFreeAsync(inputArena, 0); inputArena = nullptr; };
  try {
    const double t0 = nowMs();
    GpuBuildInput in{};
    in.nPrims = (uint32_t)n; in.nTriangles = d->n_triangles; in.nSpheres = d->n_spheres; in.nMaterials = d->n_materials;
    const size_t nt = d->n_triangles, ns = d->n_spheres;
    const size_t inputBytes = 6 * (3 * nt * 4 + 256) + (nt * 4 + 256) + (3 * ns * 4 + 256) + 2 * (ns * 4 + 256) +
                              ((size_t)n * 4 + 256) + 256;
    CU(cudaMallocAsync((void**)&inputArena, inputBytes, 0));
    size_t off = 0;
    in.triP0 = arenaCopy(d->tri_p0, 3 * nt, inputArena, off); in.triP1 = arenaCopy(d->tri_p1, 3 * nt, inputArena, off);
    in.triP2 = arenaCopy(d->tri_p2, 3 * nt, inputArena, off); in.triN0 = arenaCopy(d->tri_n0, 3 * nt, inputArena, off);
    in.triN1 = arenaCopy(d->tri_n1, 3 * nt, inputArena, off); in.triN2 = arenaCopy(d->tri_n2, 3 * nt, inputArena, off);
    in.triMaterial = arenaCopy(d->tri_material, nt, inputArena, off);
    in.sphCenter = arenaCopy(d->sph_center, 3 * ns, inputArena, off);
    in.sphRadius = arenaCopy(d->sph_radius, ns, inputArena, off);
    in.sphMaterial = arenaCopy(d->sph_material, ns, inputArena, off);
    in.primOrder = arenaCopy(d->prim_order, (size_t)n, inputArena, off);
    CU(cudaStreamSynchronize(0));
    const double t1 = nowMs();
    if (!buildBvhOnDevice(in, d->bvh_max_depth, d->split_mode, bo)) {
      cleanup(); freeGpuBuildOutput(bo);
      throw CudaFailure{bo.error, bo.where, __FILE__, __LINE__};
    }
    cleanup();
    const double t2 = nowMs();
    if (bo.errorFlags & 4u) { freeGpuBuildOutput(bo); return fail(YAHR_ERR_INVALID_ARGUMENT, "prim_order: index out of range"); }
    if (bo.errorFlags & 1u) { freeGpuBuildOutput(bo); return fail(YAHR_ERR_NON_FINITE_INPUT, "non-finite geometry"); }
    if (bo.errorFlags & 2u) {
      freeGpuBuildOutput(bo);
      return fail(YAHR_ERR_UNKNOWN_MATERIAL, "material index out of range");     // the reference: Map.! (main.hs:55)
    }
    if (bo.unsupported) { freeGpuBuildOutput(bo); return -1; }
    if (bo.tooDeep || bo.depth + 1 > YAHR_B200_MAX_STACK) {
      freeGpuBuildOutput(bo);
      return fail(YAHR_ERR_BVH_TOO_DEEP, "BVH depth " + std::to_string(bo.depth) + " exceeds the traversal stack (" +
                                             std::to_string(YAHR_B200_MAX_STACK) + ")");
    }
    sc = new yahr_scene();
    CU(cudaGetDevice(&sc->device));
    CU(cudaDeviceGetAttribute(&sc->numSMs, cudaDevAttrMultiProcessorCount, sc->device));
    sc->d_nodes = bo.flat; sc->d_prims = bo.prims; sc->d_normals = bo.normals; sc->d_multi = bo.multi;
    sc->d_order = bo.order;
    bo.flat = nullptr; bo.prims = nullptr; bo.normals = nullptr; bo.multi = nullptr; bo.order = nullptr;
    uint64_t bytes = (uint64_t)bo.nInner * 64 + (uint64_t)n * 96 + (uint64_t)bo.nMulti * 8 + (uint64_t)n * 4;
    uploadSmallTables(sc, d, bytes);
    { int rcA = uploadAreaLights(sc, d, bytes); if (rcA) { delete sc; return rcA; } }
    sc->dev.nodes = sc->d_nodes; sc->dev.prims = sc->d_prims; sc->dev.normals = sc->d_normals;
    sc->dev.multiLeaves = sc->d_multi;
    sc->dev.rootRef = n ? bo.rootRef : kDevRefNull;
    for (int c = 0; c < 3; ++c) { sc->dev.rootLo[c] = bo.rootBox[c]; sc->dev.rootHi[c] = bo.rootBox[3 + c]; }
    sc->info.n_primitives = (uint32_t)n; sc->info.n_nodes = bo.nInner; sc->info.n_multi_leaves = bo.nMulti;
    sc->nTriangles = d->n_triangles;
    sc->info.depth = bo.depth; sc->info.device_bytes = bytes;
    sc->info.build_ms = t2 - t1; sc->info.upload_ms = t1 - t0;
    sc->info.built_on_device = 1;
    attachWideTree(sc);
    CU(cudaDeviceSynchronize());
    *out = sc;
    return YAHR_OK;
  } catch (const CudaFailure& f) {
    cleanup(); freeGpuBuildOutput(bo); delete sc;
    return cudaFail(f);
  } catch (const std::bad_alloc&) {
    cleanup(); freeGpuBuildOutput(bo); delete sc;
    return fail(YAHR_ERR_OUT_OF_MEMORY, "host out of memory");
  }
}

}  // namespace

extern "C" {

int yahr_b200_abi_version(void) { return YAHR_B200_ABI_VERSION; }

int yahr_b200_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(yahr_scene_desc);
    case 1: return (int)sizeof(yahr_camera);
    case 2: return (int)sizeof(yahr_render_opts);
    case 3: return (int)sizeof(yahr_stats);
    case 4: return (int)sizeof(yahr_scene_info);
    default: return -1;
  }
}

int yahr_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* yahr_b200_last_error(void) { return g_lastError.c_str(); }

int yahr_b200_scene_create(const yahr_scene_desc* desc, yahr_scene** out) {
  if (!out) return fail(YAHR_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  yahr_scene* sc = nullptr;
  try {
    if (yahr_b200_device_count() < 1)
      return fail(YAHR_ERR_NO_DEVICE, "no CUDA device available (libyahr_b200 has no CPU fallback)");
    // Default: build the BVH on the device (both split modes).  The degenerate cases the device builder
    // declines, and YAHR_B200_HOST_BUILD=1, use the host builder below.
    if (desc && gpuBuildSupported(desc->split_mode) && !getenv("YAHR_B200_HOST_BUILD")) {
      int rcDev = createSceneOnDevice(desc, &sc);
      if (rcDev == YAHR_OK) { *out = sc; return YAHR_OK; }
      if (rcDev > 0) return rcDev;          // a real error; rcDev < 0: declined, fall through to the host builder
      sc = nullptr;
    }
    std::vector<HostPrim> prims;
    std::vector<Box> bounds;
    std::string err;
    int rc = gatherPrimitives(desc, prims, bounds, err);
    if (rc) return fail(rc, err);

    const double t0 = nowMs();
    HostBvh bvh;
    buildReferenceBvh(bounds, desc->bvh_max_depth, desc->split_mode, bvh);
    const double t1 = nowMs();
    if (bvh.maxStack > YAHR_B200_MAX_STACK)
      return fail(YAHR_ERR_BVH_TOO_DEEP, "BVH depth " + std::to_string(bvh.depth) + " exceeds the traversal stack (" +
                                             std::to_string(YAHR_B200_MAX_STACK) + ")");

    // primitives and normals in DFS leaf order
    const size_t n = prims.size();
    std::vector<float4> recs(3 * n), nrm(3 * n);
    auto bits = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
    for (size_t i = 0; i < n; ++i) {
      const uint32_t id = bvh.order[i];
      const HostPrim& p = prims[id];
      const float meta = bits((p.material << 1) | (p.kind & 1u));
      if (p.kind == 1) {
        const f3 e1 = p.b - p.a, e2 = p.c - p.a;                     // Shapes.hs:38-39
        recs[3 * i + 0] = make_float4(p.a.x, p.a.y, p.a.z, meta);
        recs[3 * i + 1] = make_float4(e1.x, e1.y, e1.z, bits(id));
        recs[3 * i + 2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
        nrm[3 * i + 0] = make_float4(p.n0.x, p.n0.y, p.n0.z, 0.0f);
        nrm[3 * i + 1] = make_float4(p.n1.x, p.n1.y, p.n1.z, 0.0f);
        nrm[3 * i + 2] = make_float4(p.n2.x, p.n2.y, p.n2.z, 0.0f);
      } else {
        recs[3 * i + 0] = make_float4(p.a.x, p.a.y, p.a.z, meta);
        recs[3 * i + 1] = make_float4(p.radius, 0.0f, 0.0f, bits(id));
        recs[3 * i + 2] = make_float4(0, 0, 0, 0);
        nrm[3 * i + 0] = nrm[3 * i + 1] = nrm[3 * i + 2] = make_float4(0, 0, 0, 0);
      }
    }
    std::vector<float4> mats(2 * (size_t)desc->n_materials), lights(2 * (size_t)desc->n_lights);
    for (uint32_t m = 0; m < desc->n_materials; ++m) {
      const float* s = desc->materials + 7 * (size_t)m;
      mats[2 * m + 0] = make_float4(s[0], s[1], s[2], s[6]);
      mats[2 * m + 1] = make_float4(s[3], s[4], s[5], 0.0f);
    }
    for (uint32_t l = 0; l < desc->n_lights; ++l) {
      const float* s = desc->lights + 6 * (size_t)l;
      lights[2 * l + 0] = make_float4(s[0], s[1], s[2], 0.0f);
      lights[2 * l + 1] = make_float4(s[3], s[4], s[5], 0.0f);
    }
    std::vector<uint2> multi(bvh.multiLeaves.size() / 2);
    for (size_t k = 0; k < multi.size(); ++k) multi[k] = make_uint2(bvh.multiLeaves[2 * k], bvh.multiLeaves[2 * k + 1]);

    sc = new yahr_scene();
    CU(cudaGetDevice(&sc->device));
    CU(cudaDeviceGetAttribute(&sc->numSMs, cudaDevAttrMultiProcessorCount, sc->device));
    const double t2 = nowMs();
    uint64_t bytes = 0;
    static_assert(sizeof(FlatNode) == 4 * sizeof(float4), "node layout");
    if (!bvh.flat.empty()) {
      CU(cudaMalloc(&sc->d_nodes, bvh.flat.size() * sizeof(FlatNode)));
      CU(cudaMemcpy(sc->d_nodes, bvh.flat.data(), bvh.flat.size() * sizeof(FlatNode), cudaMemcpyHostToDevice));
      bytes += bvh.flat.size() * sizeof(FlatNode);
    }
    sc->d_prims = devUpload(recs, bytes);
    sc->d_normals = devUpload(nrm, bytes);
    sc->d_multi = devUpload(multi, bytes);
    sc->d_materials = devUpload(mats, bytes);
    sc->d_lights = devUpload(lights, bytes);
    sc->d_order = devUpload(bvh.order, bytes);
    CU(cudaMalloc(&sc->d_counters, 8 * sizeof(unsigned long long)));
    CU(cudaEventCreate(&sc->ev0));
    CU(cudaEventCreate(&sc->ev1));
    for (auto& e : sc->phaseEv) CU(cudaEventCreate(&e));
    CU(cudaDeviceSynchronize());
    const double t3 = nowMs();

    sc->dev.nodes = sc->d_nodes; sc->dev.prims = sc->d_prims; sc->dev.normals = sc->d_normals;
    sc->dev.multiLeaves = sc->d_multi; sc->dev.materials = sc->d_materials; sc->dev.lights = sc->d_lights;
    sc->dev.nLights = desc->n_lights;
    { int rcA = uploadAreaLights(sc, desc, bytes); if (rcA) { delete sc; return rcA; } }
    sc->dev.rootRef = bvh.rootRef;
    sc->dev.rootLo[0] = bvh.rootBox.lo.x; sc->dev.rootLo[1] = bvh.rootBox.lo.y; sc->dev.rootLo[2] = bvh.rootBox.lo.z;
    sc->dev.rootHi[0] = bvh.rootBox.hi.x; sc->dev.rootHi[1] = bvh.rootBox.hi.y; sc->dev.rootHi[2] = bvh.rootBox.hi.z;
    sc->info.n_primitives = (uint32_t)n;
    sc->nTriangles = desc->n_triangles;
    sc->info.n_nodes = (uint32_t)bvh.flat.size();
    sc->info.n_multi_leaves = (uint32_t)multi.size();
    sc->info.depth = bvh.depth;
    sc->info.device_bytes = bytes;
    sc->info.build_ms = t1 - t0;
    sc->info.upload_ms = t3 - t2;
    sc->info.built_on_device = 0;
    attachWideTree(sc);
    *out = sc;
    return YAHR_OK;
  } catch (const CudaFailure& f) {
    delete sc;
    return cudaFail(f);
  } catch (const std::bad_alloc&) {
    delete sc;
    return fail(YAHR_ERR_OUT_OF_MEMORY, "host out of memory");
  } catch (const std::exception& e) {
    delete sc;
    return fail(YAHR_ERR_INTERNAL, e.what());
  }
}

void yahr_b200_scene_destroy(yahr_scene* scene) {
  if (!scene) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(scene->device);
  delete scene;
  cudaSetDevice(prev);
}

// Inspection: download the BVH that lives on the device (used to check the GPU builder against the
// host builder).  Any output may be NULL.  nodes_out: n_nodes x 16 floats (the 64-byte node layout),
// multi_out: n_multi_leaves x 2 uint32, order_out: n_primitives uint32.
int yahr_b200_scene_download_bvh(const yahr_scene* scene, uint32_t* order_out, float* nodes_out, uint32_t* multi_out,
                                 uint32_t* root_ref_out, float root_box_out[6]) {
  if (!scene) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  try {
    const yahr_scene_info& i = scene->info;
    if (order_out && i.n_primitives)
      CU(cudaMemcpy(order_out, scene->d_order, (size_t)i.n_primitives * 4, cudaMemcpyDeviceToHost));
    if (nodes_out && i.n_nodes) CU(cudaMemcpy(nodes_out, scene->d_nodes, (size_t)i.n_nodes * 64, cudaMemcpyDeviceToHost));
    if (multi_out && i.n_multi_leaves)
      CU(cudaMemcpy(multi_out, scene->d_multi, (size_t)i.n_multi_leaves * 8, cudaMemcpyDeviceToHost));
    if (root_ref_out) *root_ref_out = scene->dev.rootRef;
    if (root_box_out)
      for (int c = 0; c < 3; ++c) { root_box_out[c] = scene->dev.rootLo[c]; root_box_out[3 + c] = scene->dev.rootHi[c]; }
    return YAHR_OK;
  } catch (const CudaFailure& f) {
    return cudaFail(f);
  }
}

int yahr_b200_scene_download_wide(const yahr_scene* scene, float* wide_out) {
  if (!scene || !wide_out) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  try {
    if (scene->info.n_wide_nodes)
      CU(cudaMemcpy(wide_out, scene->d_wide, (size_t)scene->info.n_wide_nodes * kWideNodeVec * sizeof(float4),
                    cudaMemcpyDeviceToHost));
    return YAHR_OK;
  } catch (const CudaFailure& f) {
    return cudaFail(f);
  }
}

int yahr_b200_scene_download_compressed(const yahr_scene* scene, float* cwide_out, float* leaf_box_out, float* multi_box_out) {
  if (!scene) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!scene->d_cwide) return fail(YAHR_ERR_INVALID_ARGUMENT, "the scene has no compressed wide nodes");
  try {
    const yahr_scene_info& i = scene->info;
    if (cwide_out && i.n_wide_nodes)
      CU(cudaMemcpy(cwide_out, scene->d_cwide, (size_t)i.n_wide_nodes * 4 * sizeof(float4), cudaMemcpyDeviceToHost));
    if (leaf_box_out && i.n_primitives)
      CU(cudaMemcpy(leaf_box_out, scene->d_leafBox, (size_t)i.n_primitives * 2 * sizeof(float4), cudaMemcpyDeviceToHost));
    if (multi_box_out && i.n_multi_leaves)
      CU(cudaMemcpy(multi_box_out, scene->d_multiBox, (size_t)i.n_multi_leaves * 2 * sizeof(float4), cudaMemcpyDeviceToHost));
    return YAHR_OK;
  } catch (const CudaFailure& f) {
    return cudaFail(f);
  }
}

int yahr_b200_scene_info(const yahr_scene* scene, yahr_scene_info* out) {
  if (!scene || !out) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = scene->info;
  return YAHR_OK;
}

int yahr_b200_render_device(yahr_scene* scene, const yahr_camera* cam, const yahr_render_opts* opts, float* d_rgb,
                            uint32_t* d_primid, void* stream, yahr_stats* stats) {
  if (!scene || !cam || !opts || !d_rgb) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  try {
    return renderCommon(scene, cam, opts, d_rgb, d_primid, (cudaStream_t)stream, stats, false);
  } catch (const CudaFailure& f) {
    return cudaFail(f);
  } catch (const std::exception& e) {
    return fail(YAHR_ERR_INTERNAL, e.what());
  }
}

// The same frame through the COUNTING build of the wavefront kernels (wavefront_count.cu): identical results, plus the
// GPU's own work counters -- counts_out[0..7] for the closest-hit walks, [8..15] for the any-hit walks, each
// {4-wide node visits, binary node visits, primitive tests, normal fetches, stack pushes, stack pops, -, shaded hits}.
int yahr_b200_render_device_counted(yahr_scene* scene, const yahr_camera* cam, const yahr_render_opts* opts, float* d_rgb,
                                    uint32_t* d_primid, void* stream, yahr_stats* stats, uint64_t counts_out[16]) {
  if (!scene || !cam || !opts || !d_rgb || !counts_out) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  try {
    return renderCommon(scene, cam, opts, d_rgb, d_primid, (cudaStream_t)stream, stats, false, counts_out);
  } catch (const CudaFailure& f) {
    return cudaFail(f);
  } catch (const std::exception& e) {
    return fail(YAHR_ERR_INTERNAL, e.what());
  }
}

static void renderStreamedRows(yahr_scene* scene, FramePlan& plan, float* rgb_out, unsigned char* rgb8_out,
                               uint32_t* primid_out, yahr_stats* stats, double w0, double copyBytesPerMs);

// Device-buffer multi-GPU entry: render the tile rows of this shard into the LOCAL frame, then push exactly those
// pixel rows into the gather frame (usually rank 0's, mapped with yahr_b200_ipc_open) with device-to-device copies on
// the same stream: a few large NVLink transfers instead of one small remote store per pixel.
int yahr_b200_render_device_shard(yahr_scene* scene, const yahr_camera* cam, const yahr_render_opts* opts_in,
                                  int shard_index, int shard_count, float* d_rgb_local, float* d_rgb_gather,
                                  uint32_t* d_primid_local, uint32_t* d_primid_gather, void* stream, yahr_stats* stats) {
  if (!scene || !cam || !opts_in || !d_rgb_local) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count)
    return fail(YAHR_ERR_INVALID_ARGUMENT, "shard_index / shard_count invalid");
  try {
    yahr_render_opts o = *opts_in;
    o.tile_stride = shard_count; o.tile_offset = shard_index; o.reserved[1] = 1;      // whole rows of the tile grid
    cudaStream_t st = (cudaStream_t)stream;
    // STREAMED push (one light slot, 1 spp, a gather frame that is not the local one): ONE launch of the per-batch
    // kernel; every finished tile row is copied into the gather frame by the copy engine while the rest of the share
    // is traced (the mechanism of the host-buffer entry, renderStreamedRows, with a peer frame as the destination), so
    // the exchange overlaps the rendering instead of following it.  The call then returns when this shard's rows ARE in
    // the gather frame (it blocks the host thread; work enqueued on `stream` before / after it is ordered before / after).
    // YAHR_B200_SHARD_STREAM=0 keeps the render-then-push path below.
    {
      const bool push = (d_rgb_gather && d_rgb_gather != d_rgb_local);
      const char* env = getenv("YAHR_B200_SHARD_STREAM");
      if (push && !(env && atoi(env) == 0) && o.spp == 1 && !stats) {
        FramePlan plan;
        int prc = planFrame(scene, cam, &o, d_rgb_local, d_primid_local, plan);
        if (prc) return prc;
        const TileSet& pts = *plan.ts;
        const bool pushPid0 = d_primid_gather && d_primid_local && d_primid_gather != d_primid_local;
        if (plan.wavefront && !plan.W.dense && !plan.W.twoSlot && pts.rowY.size() >= 2 && pts.d_rowOfV && scene->dev.wide &&
            scene->dev.nAreaLights == 0 && plan.W.wideTree && o.traversal == YAHR_TRAVERSAL_REFERENCE) {
          for (auto& s2 : scene->renderStream) if (!s2) CU(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
          if (!scene->copyStream) CU(cudaStreamCreateWithFlags(&scene->copyStream, cudaStreamNonBlocking));
          if (!scene->evIn) CU(cudaEventCreateWithFlags(&scene->evIn, cudaEventDisableTiming));
          CU(cudaEventRecord(scene->evIn, st));
          CU(cudaStreamWaitEvent(scene->renderStream[0], scene->evIn, 0));
          renderStreamedRows(scene, plan, d_rgb_gather, nullptr, pushPid0 ? d_primid_gather : nullptr, nullptr, nowMs(),
                             400.0e6);
          return YAHR_OK;                      // both internal streams have been synchronised: nothing left in flight
        }
      }
    }
    int rc = renderCommon(scene, cam, &o, d_rgb_local, d_primid_local, st, nullptr, false);
    if (rc) return rc;
    CameraSetup cs;
    std::string err;
    rc = setupCamera(cam, cs, err);
    if (rc) return fail(rc, err);
    const TileSet& ts = tilesFor(scene, cs.width, cs.height, shard_count, shard_index, 1);
    const bool pushRgb = d_rgb_gather && d_rgb_gather != d_rgb_local;
    const bool pushPid = d_primid_gather && d_primid_local && d_primid_gather != d_primid_local;
    const uint32_t nRows = (uint32_t)ts.rowY.size();
    // equally tall, equally spaced rows (the usual case: 2160 = 135 x 16): ONE strided copy per buffer
    bool uniform = nRows >= 2;
    for (uint32_t k = 1; uniform && k < nRows; ++k)
      uniform = (ts.rowY[k].y - ts.rowY[k].x) == (ts.rowY[0].y - ts.rowY[0].x) &&
                (ts.rowY[k].x - ts.rowY[k - 1].x) == (ts.rowY[1].x - ts.rowY[0].x);
    if ((pushRgb || pushPid) && uniform) {
      const size_t y0 = (size_t)ts.rowY[0].x, rowsTall = (size_t)(ts.rowY[0].y - ts.rowY[0].x);
      const size_t spacing = (size_t)(ts.rowY[1].x - ts.rowY[0].x);
      if (pushRgb) {
        const size_t rowBytes = (size_t)cs.width * 3 * sizeof(float);
        CU(cudaMemcpy2DAsync((char*)d_rgb_gather + y0 * rowBytes, spacing * rowBytes, (const char*)d_rgb_local + y0 * rowBytes,
                             spacing * rowBytes, rowsTall * rowBytes, nRows, cudaMemcpyDeviceToDevice, st));
      }
      if (pushPid) {
        const size_t rowBytes = (size_t)cs.width * sizeof(uint32_t);
        CU(cudaMemcpy2DAsync((char*)d_primid_gather + y0 * rowBytes, spacing * rowBytes,
                             (const char*)d_primid_local + y0 * rowBytes, spacing * rowBytes, rowsTall * rowBytes, nRows,
                             cudaMemcpyDeviceToDevice, st));
      }
    }
    uint32_t r = 0;
    while ((pushRgb || pushPid) && !uniform && r < nRows) {
      uint32_t e = r + 1;
      while (e < nRows && ts.rowY[e].x == ts.rowY[e - 1].y) ++e;
      const size_t y0 = (size_t)ts.rowY[r].x, y1 = (size_t)ts.rowY[e - 1].y;
      if (pushRgb) {
        const size_t rowBytes = (size_t)cs.width * 3 * sizeof(float);
        CU(cudaMemcpyAsync((char*)d_rgb_gather + y0 * rowBytes, (const char*)d_rgb_local + y0 * rowBytes, (y1 - y0) * rowBytes,
                           cudaMemcpyDeviceToDevice, st));
      }
      if (pushPid) {
        const size_t rowBytes = (size_t)cs.width * sizeof(uint32_t);
        CU(cudaMemcpyAsync((char*)d_primid_gather + y0 * rowBytes, (const char*)d_primid_local + y0 * rowBytes,
                           (y1 - y0) * rowBytes, cudaMemcpyDeviceToDevice, st));
      }
      r = e;
    }
    if (stats) {
      std::memset(stats, 0, sizeof(*stats));
      CU(cudaStreamSynchronize(st));
      unsigned long long c[3];
      CU(cudaMemcpy(c, scene->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
      stats->n_primary = c[0]; stats->n_shadow = c[1]; stats->n_secondary = c[2];
      stats->tiles = ts.n;
    }
    return YAHR_OK;
  } catch (const CudaFailure& f) {
    return cudaFail(f);
  } catch (const std::exception& e) {
    return fail(YAHR_ERR_INTERNAL, e.what());
  }
}

// Host-buffer entry, STREAMED ROWS (one light slot, 1 spp, float frame -- the reference's own configuration).  The
// whole share is ONE launch of k_wf_fused; the kernel counts finalised pixels per row of the tile grid and publishes a
// completed row in mapped host memory (wavefront.cu, rowsSignal); this thread polls the flags and starts the
// device-to-host copy of every completed run of rows at once, so the copy engine runs under the traversal at row
// granularity and the persistent kernel pays its ramp-up and tail once (bands: once per band).  If the kernel ends
// before every flag has been seen the remaining rows are simply copied then: the accounting can never lose a row.
static void renderStreamedRows(yahr_scene* scene, FramePlan& plan, float* rgb_out, unsigned char* rgb8_out,
                               uint32_t* primid_out, yahr_stats* stats, double w0, double copyBytesPerMs) {
  const TileSet& ts = *plan.ts;
  const uint32_t nRowsS = (uint32_t)ts.rowY.size();
  const int W_ = plan.cs.width;
  cudaStream_t rs = scene->renderStream[0], cp = scene->copyStream;
  static const bool timelineS = getenv("YAHR_B200_TIMELINE") != nullptr;
  if (nRowsS > scene->rowCap) {
    cudaFree(scene->d_rowDone); scene->d_rowDone = nullptr;
    if (scene->h_rowFlags) cudaFreeHost(scene->h_rowFlags);
    scene->h_rowFlags = nullptr; scene->rowCap = 0;
    CU(cudaMalloc(&scene->d_rowDone, nRowsS * sizeof(uint32_t)));
    CU(cudaHostAlloc((void**)&scene->h_rowFlags, nRowsS * sizeof(uint32_t), cudaHostAllocMapped));
    std::memset(scene->h_rowFlags, 0, nRowsS * sizeof(uint32_t));
    CU(cudaHostGetDevicePointer((void**)&scene->d_rowFlags, scene->h_rowFlags, 0));
    scene->rowCap = nRowsS; scene->rowSeq = 0;
  }
  if (++scene->rowSeq == 0u) {                                    // sequence wrapped: start over with clean flags
    std::memset(scene->h_rowFlags, 0, scene->rowCap * sizeof(uint32_t));
    scene->rowSeq = 1u;
  }
  const uint32_t seq = scene->rowSeq;
  plan.W.rowOfV = ts.d_rowOfV; plan.W.rowItems = ts.d_rowItems; plan.W.rowDone = scene->d_rowDone;
  plan.W.rowFlags = scene->d_rowFlags; plan.W.rowSeq = seq;
  // 8-bit entry: the kernel applies the reference's output stage (main.hs:142) to every pixel it finalises and stores
  // three bytes instead of three floats, so a finished row is ready for its (4x smaller) copy as it stands
  plan.W.rgb8 = rgb8_out ? scene->d_rgb8 : nullptr;
  // rows complete in item order only when every batch is final at once: the fused kernel (with the two-kernel set
  // every lit row completes in the shadow phase, after the whole primary trace)
  // k_wf_fused by default (C4 1.59 ms against 1.75 - 1.90 for k_wf_persist on a whole frame, profiles/r2c-r2e);
  // YAHR_B200_HOST_FUSED=3 selects k_wf_persist, 2 the 72-register build of k_wf_fused, 0 the two-kernel set
  // A large 8-bit frame streams from the TWO-kernel set: its rows only become final in the shadow phase, but 24.9 MB (C4)
  // need 0.44 ms of copy and the shadow kernel alone runs 0.43 ms, so the faster kernels decide the call (profiles/r2o).
  {
    const char* f = getenv("YAHR_B200_HOST_FUSED");
    // (8-bit: by size like the device-resident frames -- plan.W.fused as planFrame set it; float: always per batch)
    const uint32_t k = f ? (uint32_t)atoi(f) & 3u : (rgb8_out && plan.P.depth == 1 ? (plan.W.fused ? 1u : 0u) : 1u);
    plan.W.persist = (k == 3u && plan.P.depth == 1) ? 1u : 0u;
    plan.W.fused = k == 3u ? 1u : k;
  }
  uint32_t launches = 0;
  CU(cudaMemsetAsync(scene->d_rowDone, 0, nRowsS * sizeof(uint32_t), rs));
  CU(cudaMemsetAsync(scene->d_counters, 0, 3 * sizeof(unsigned long long), rs));
  CU(cudaEventRecord(scene->ev0, rs));
  enqueueTiles(scene, plan, 0, ts.n, rs, &launches, scene->phaseEv);
  CU(cudaEventRecord(scene->ev1, rs));
  uint64_t d2h = 0;
  uint32_t nCopies = 0;
  const size_t rowBytes = (size_t)W_ * 3 * (rgb8_out ? 1 : sizeof(float)), idBytes = (size_t)W_ * sizeof(uint32_t);
  char* const hostFrame = rgb8_out ? (char*)rgb8_out : (char*)rgb_out;
  // the frame the kernels of this plan render into (the library's own frame for the host-buffer entry, the caller's
  // local frame for the device shard entry); destinations may be host memory or a peer GPU's frame (UVA decides)
  const char* const devFrame = rgb8_out ? (const char*)scene->d_rgb8 : (const char*)plan.W.base.rgb;
  auto copyRows = [&](uint32_t r, uint32_t e) {
    const int y0 = ts.rowY[r].x, y1 = ts.rowY[e - 1].y;
    CU(cudaMemcpyAsync(hostFrame + (size_t)y0 * rowBytes, devFrame + (size_t)y0 * rowBytes,
                       (size_t)(y1 - y0) * rowBytes, cudaMemcpyDefault, cp));
    d2h += (uint64_t)(y1 - y0) * rowBytes;
    if (primid_out) {
      CU(cudaMemcpyAsync((char*)primid_out + (size_t)y0 * idBytes, (const char*)plan.W.base.primid + (size_t)y0 * idBytes,
                         (size_t)(y1 - y0) * idBytes, cudaMemcpyDefault, cp));
      d2h += (uint64_t)(y1 - y0) * idBytes;
    }
    ++nCopies;
  };
  // Copy issue policy.  Every copy costs ~5 us of copy-engine idle time whatever its size, so while the engine
  // has work queued a short run of finished rows waits for its neighbours (rows finish roughly in order) and
  // the copies grow exactly when the engine is the bottleneck; when it is about to run dry, whatever is ready
  // goes out at once.  The engine's backlog is MODELLED (bytes queued at 48 GB/s + 8 us per copy), not queried: events
  // between the copies would serialise them.  Runs end at the gaps between a shard's blocks of rows.
  const volatile uint32_t* flags = scene->h_rowFlags;
  std::vector<unsigned char> issued(nRowsS, 0);
  uint32_t nIssued = 0, lowest = 0, spins = 0, flaggedRows = 0;
  // a run that has grown to this size goes out even when the engine has work
  const uint64_t maxHeldBytes = getenv("YAHR_B200_STREAM_MAX_HELD_KB")
                                    ? (uint64_t)atoi(getenv("YAHR_B200_STREAM_MAX_HELD_KB")) << 10 : (uint64_t)6 << 20;
  double busyUntil = 0.0;                                         // ms on the nowMs() clock
  bool kernelsDone = false;
  // whatever goes wrong below, the persistent kernel (still storing to the mapped flags) and the queued copies into the
  // caller's buffers must have drained before the error leaves this call
  struct Drain {
    cudaStream_t a, b; bool armed;
    ~Drain() { if (armed) { cudaStreamSynchronize(a); cudaStreamSynchronize(b); cudaGetLastError(); } }
  } drain{rs, cp, true};
  while (nIssued < nRowsS) {
    bool progress = false;
    while (lowest < nRowsS && issued[lowest]) ++lowest;
    for (uint32_t r = lowest; r < nRowsS;) {
      if (issued[r] || !(kernelsDone || flags[r] == seq)) { ++r; continue; }
      uint32_t e = r + 1;
      while (e < nRowsS && !issued[e] && (kernelsDone || flags[e] == seq) && ts.rowY[e].x == ts.rowY[e - 1].y) ++e;
      const uint64_t bytes = (uint64_t)(ts.rowY[e - 1].y - ts.rowY[r].x) * rowBytes;
      const bool canGrow = e < nRowsS && !issued[e] && ts.rowY[e].x == ts.rowY[e - 1].y;     // row e is not finished yet
      const double now = nowMs();
      if (!kernelsDone && canGrow && bytes < maxHeldBytes && busyUntil - now > 0.05) { r = e; continue; }
      copyRows(r, e);
      busyUntil = (busyUntil > now ? busyUntil : now) + (double)bytes / copyBytesPerMs + 0.008;    // measured under load
      if (!kernelsDone) flaggedRows += e - r;
      for (uint32_t k = r; k < e; ++k) issued[k] = 1;
      nIssued += e - r;
      progress = true;
      r = e;
    }
    if (!progress && !kernelsDone && (++spins & 15u) == 0u) {
      const cudaError_t q = cudaStreamQuery(rs);
      if (q == cudaSuccess) kernelsDone = true;                  // everything is rendered: copy what is left
      else if (q != cudaErrorNotReady) throw CudaFailure{q, "cudaStreamQuery(render stream)", __FILE__, __LINE__};
    }
  }
  cudaEvent_t copyEnd = nullptr;
  if (timelineS) { CU(cudaEventCreate(&copyEnd)); CU(cudaEventRecord(copyEnd, cp)); }
  unsigned long long c[3];
  CU(cudaMemcpyAsync(c, scene->d_counters, sizeof(c), cudaMemcpyDeviceToHost, rs));
  CU(cudaStreamSynchronize(rs));
  CU(cudaStreamSynchronize(cp));
  drain.armed = false;
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, scene->ev0, scene->ev1));
  scene->lastHostGpuMs = ms;
  if (timelineS) {
    float tc = 0;
    CU(cudaEventElapsedTime(&tc, scene->ev0, copyEnd));
    cudaEventDestroy(copyEnd);
    fprintf(stderr, "[yahr_b200 timeline] streamed rows: wall %.3f ms; render %.3f copy end %.3f; %u rows, %u flagged "
            "before the kernels ended, %u copies\n", nowMs() - w0, ms, tc, nRowsS, flaggedRows, nCopies);
  }
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->n_primary = c[0]; stats->n_shadow = c[1]; stats->n_secondary = c[2];
    stats->gpu_ms = ms;
    if (ts.n) {
      for (int k = 0; k < 3; ++k) {
        float pm = 0;
        CU(cudaEventElapsedTime(&pm, scene->phaseEv[k], scene->phaseEv[k + 1]));
        stats->phase_ms[k] = pm;
      }
    }
    stats->launches = launches;
    stats->tiles = ts.n;
    stats->h2d_bytes = sizeof(WavefrontParams) * (uint64_t)launches;
    stats->d2h_bytes = d2h;
    stats->wall_ms = nowMs() - w0;
  }
}

static int renderHost(yahr_scene* scene, const yahr_camera* cam, int recursion_depth, int spp, uint64_t seed,
                      float* rgb_out, unsigned char* rgb8_out, uint32_t* primid_out, yahr_stats* stats,
                      int shardIndex = 0, int shardCount = 1) {
  if (!scene || !cam || (!rgb_out && !rgb8_out)) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  try {
    const double w0 = nowMs();
    yahr_render_opts o{};
    o.recursion_depth = recursion_depth; o.spp = spp; o.seed = seed;
    if (shardCount < 1 || shardIndex < 0 || shardIndex >= shardCount)
      return fail(YAHR_ERR_INVALID_ARGUMENT, "shard_index / shard_count invalid");
    // the frame is cut into whole rows of the reference's tile grid, dealt to the shards in blocks of up to four
    // consecutive rows (tilesFor, byRows = 2); this call renders the blocks shardIndex, shardIndex + shardCount, ...
    // (every row for the single-GPU entry)
    o.traversal = YAHR_TRAVERSAL_REFERENCE; o.tile_stride = shardCount; o.tile_offset = shardIndex; o.reserved[1] = 2;
    {
      CameraSetup cs;
      std::string err;
      int rc0 = setupCamera(cam, cs, err);
      if (rc0) return fail(rc0, err);
      const size_t px0 = (size_t)cs.width * cs.height;
      if (px0 > scene->framePixels) {
        cudaFree(scene->d_rgb); cudaFree(scene->d_primid);
        scene->d_rgb = nullptr; scene->d_primid = nullptr; scene->framePixels = 0;
        CU(cudaMalloc(&scene->d_rgb, px0 * 3 * sizeof(float)));
        CU(cudaMalloc(&scene->d_primid, px0 * sizeof(uint32_t)));
        scene->framePixels = px0;
      }
      if (rgb8_out && px0 > scene->frame8Pixels) {
        cudaFree(scene->d_rgb8); scene->d_rgb8 = nullptr; scene->frame8Pixels = 0;
        CU(cudaMalloc(&scene->d_rgb8, px0 * 3));
        scene->frame8Pixels = px0;
      }
    }
    FramePlan plan;
    int rc = planFrame(scene, cam, &o, scene->d_rgb, primid_out ? scene->d_primid : nullptr, plan);
    if (rc) return rc;
    const TileSet& ts = *plan.ts;
    const int W_ = plan.cs.width, H_ = plan.cs.height;
    for (auto& st : scene->renderStream) if (!st) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (!scene->copyStream) CU(cudaStreamCreateWithFlags(&scene->copyStream, cudaStreamNonBlocking));
    cudaStream_t rs = scene->renderStream[0], rs1 = scene->renderStream[1], cp = scene->copyStream;
    // the scratch queues are shared with yahr_b200_render_device calls that may still be in flight
    // on a caller stream
    CU(cudaDeviceSynchronize());
    // test hook: the device frame persists between calls, so a row copied before its pixels have landed would go
    // unnoticed when the same frame is rendered twice; YAHR_B200_POISON_FRAME=1 fills it with NaN first
    if (getenv("YAHR_B200_POISON_FRAME")) {
      CU(cudaMemset(scene->d_rgb, 0xFF, (size_t)W_ * H_ * 3 * sizeof(float)));
      if (primid_out) CU(cudaMemset(scene->d_primid, 0xEE, (size_t)W_ * H_ * sizeof(uint32_t)));
    }

    // Output strategy: the copy-engine bands below, or the streamed rows (renderStreamedRows).  A STATIC rule: one light
    // slot at 1 spp (the reference's own configuration, float or 8-bit frame) streams its rows from the very first call
    // -- the reference's usage is one frame per process (main.hs:112-142), so a choice that needs warm-up calls would
    // never reach it, and a C ABI whose first calls behave differently is not repeatable.  Everything else uses the
    // bands.  YAHR_B200_HOST_STREAM=0 / 1 pins the bands / the streamed rows; YAHR_B200_HOST_MEASURE=1 brings back the
    // measured choice of round 1 (three calls each, then the faster one; HostStrategy) for experiments.
    HostStrategy* strategy = nullptr;
    {
      const char* env = getenv("YAHR_B200_HOST_STREAM");
      const uint32_t nRowsS = (uint32_t)ts.rowY.size();
      bool useStream = plan.wavefront && !plan.W.dense && !plan.W.twoSlot && spp == 1 && nRowsS >= 2 && ts.d_rowOfV;
      // the 8-bit frame is written by the one-slot kernels without area lights (4-wide tree)
      if (rgb8_out && !(scene->dev.wide && scene->dev.nAreaLights == 0)) useStream = false;
      if (useStream && env) useStream = atoi(env) != 0;
      else if (useStream && getenv("YAHR_B200_HOST_MEASURE")) {
        strategy = &scene->hostStrategy[std::make_tuple(W_, H_, shardIndex, shardCount, primid_out ? 1 : (rgb8_out ? 2 : 0))];
        useStream = strategy->wantStream();
      }
      if (useStream) {
        renderStreamedRows(scene, plan, rgb_out, rgb8_out, primid_out, stats, w0, 48.0e6);
        if (strategy) strategy->record(true, nowMs() - w0);
        return YAHR_OK;
      }
    }

    // The rows of this call are rendered in BANDS of consecutive owned tile rows, and every finished band is copied
    // to the caller's buffer on a copy stream while later bands render, so the device-to-host transfer overlaps
    // traversal.  Bands alternate between two render streams (each with its own shadow queue), so the shadow trace
    // of one band overlaps the primary trace of the next and the tails of the persistent kernels are filled.
    // spp > 1 accumulates per pixel across sample passes inside a band, which works the same way.
    const uint32_t nRows = (uint32_t)ts.rowY.size();
    uint32_t nBands = 1;
    if (nRows >= 2) {
      const char* env = getenv("YAHR_B200_BANDS");
      uint32_t want = env ? (uint32_t)atoi(env) : 8u;
      if (shardCount > 1 && !env) want = (8u + (uint32_t)shardCount - 1u) / (uint32_t)shardCount + 1u;   // smaller shards, fewer bands
      if (want < 1) want = 1;
      nBands = want < nRows ? want : nRows;
    }
    static const bool timeline = getenv("YAHR_B200_TIMELINE") != nullptr;     // debug: per-band completion times on stderr
    while (scene->bandEvents.size() < nBands) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, timeline ? cudaEventDefault : cudaEventDisableTiming));
      scene->bandEvents.push_back(e);
    }
    std::vector<cudaEvent_t> copyDone;
    if (timeline)
      for (uint32_t b = 0; b < nBands; ++b) { cudaEvent_t e; CU(cudaEventCreate(&e)); copyDone.push_back(e); }
    uint32_t launches = 0;
    // Band ORDER: a band costs its rendering on the GPU and then its copy on the copy engine (a two-stage flow
    // shop; the copies cost the same for every band).  When the copies dominate, cheap bands -- sky, ground beyond the
    // model -- should go first, so that their copies run under the rendering of the expensive ones (Johnson's rule).
    // The cost estimate is the number of shadow probes each band emitted in the previous call with the same band
    // layout; the first call goes top to bottom.  Order only: every band is rendered and copied either way.
    if (!scene->d_bandProbes) CU(cudaMalloc(&scene->d_bandProbes, 64 * sizeof(uint32_t)));
    if (nBands > 64) nBands = 64;
    const auto bandKey = std::make_tuple(W_, H_, (int)nBands, shardIndex, shardCount);
    std::vector<uint32_t> order(nBands);
    for (uint32_t b = 0; b < nBands; ++b) order[b] = b;
    if (plan.wavefront && scene->bandKey == bandKey && scene->bandProbes.size() == nBands && scene->lastHostGpuMs > 0.0f &&
        nRows && !getenv("YAHR_B200_BANDS_IN_ORDER")) {
      // Only when the copies are the bottleneck of the call (frame bytes at ~55 GB/s against the GPU time of the
      // previous call); a render-bound call gains nothing from reordering and keeps top to bottom.
      const double bytesPerPixel = (rgb_out ? 12.0 : 0.0) + (rgb8_out ? 3.0 : 0.0) + (primid_out ? 4.0 : 0.0);
      const double copyMs = (double)ts.nItems * bytesPerPixel / 55.0e6;
      if (copyMs >= scene->lastHostGpuMs * 0.9)
        std::stable_sort(order.begin(), order.end(),
                         [&](uint32_t a, uint32_t b) { return scene->bandProbes[a] < scene->bandProbes[b]; });
    }
    CU(cudaMemsetAsync(scene->d_bandProbes, 0, 64 * sizeof(uint32_t), rs));
    CU(cudaMemsetAsync(scene->d_counters, 0, 3 * sizeof(unsigned long long), rs));
    CU(cudaEventRecord(scene->ev0, rs));
    if (nBands > 1) CU(cudaStreamWaitEvent(rs1, scene->ev0, 0));     // counters are cleared before any band
    uint64_t d2h = 0;
    for (uint32_t pos = 0; pos < nBands && nRows; ++pos) {
      const uint32_t b = order[pos];
      const uint32_t r0 = (uint32_t)((uint64_t)nRows * b / nBands), r1 = (uint32_t)((uint64_t)nRows * (b + 1) / nBands);
      const uint32_t first = ts.rowFirst[r0], count = ts.rowFirst[r1] - ts.rowFirst[r0];
      cudaStream_t bs = (pos & 1u) ? rs1 : rs;
      enqueueTiles(scene, plan, first, count, bs, &launches, (pos == 0) ? scene->phaseEv : nullptr, (int)(pos & 1u),
                   scene->d_bandProbes + b);
      // copies: one per run of pixel rows that is contiguous in the frame (the whole band when every row is owned)
      uint32_t r = r0;
      while (r < r1) {
        uint32_t e = r + 1;
        while (e < r1 && ts.rowY[e].x == ts.rowY[e - 1].y) ++e;
        const int y0 = ts.rowY[r].x, y1 = ts.rowY[e - 1].y;
        if (rgb8_out) {                                   // output stage on the GPU: 1 byte per channel goes down
          const size_t rowElems = (size_t)W_ * 3;
          CU(launchQuantizeRgb8(scene->d_rgb, scene->d_rgb8, (size_t)y0 * rowElems, (size_t)(y1 - y0) * rowElems, bs,
                                &launches));
        }
        r = e;
      }
      CU(cudaEventRecord(scene->bandEvents[b], bs));
      CU(cudaStreamWaitEvent(cp, scene->bandEvents[b], 0));
      r = r0;
      while (r < r1) {
        uint32_t e = r + 1;
        while (e < r1 && ts.rowY[e].x == ts.rowY[e - 1].y) ++e;
        const int y0 = ts.rowY[r].x, y1 = ts.rowY[e - 1].y;
        if (rgb_out) {
          const size_t rowBytes = (size_t)W_ * 3 * sizeof(float);
          CU(cudaMemcpyAsync((char*)rgb_out + (size_t)y0 * rowBytes, (const char*)scene->d_rgb + (size_t)y0 * rowBytes,
                             (size_t)(y1 - y0) * rowBytes, cudaMemcpyDeviceToHost, cp));
          d2h += (uint64_t)(y1 - y0) * rowBytes;
        }
        if (rgb8_out) {
          const size_t rowBytes = (size_t)W_ * 3;
          CU(cudaMemcpyAsync(rgb8_out + (size_t)y0 * rowBytes, scene->d_rgb8 + (size_t)y0 * rowBytes,
                             (size_t)(y1 - y0) * rowBytes, cudaMemcpyDeviceToHost, cp));
          d2h += (uint64_t)(y1 - y0) * rowBytes;
        }
        if (primid_out) {
          const size_t idBytes = (size_t)W_ * sizeof(uint32_t);
          CU(cudaMemcpyAsync((char*)primid_out + (size_t)y0 * idBytes, (const char*)scene->d_primid + (size_t)y0 * idBytes,
                             (size_t)(y1 - y0) * idBytes, cudaMemcpyDeviceToHost, cp));
          d2h += (uint64_t)(y1 - y0) * idBytes;
        }
        r = e;
      }
      if (timeline) CU(cudaEventRecord(copyDone[b], cp));
    }
    if (nBands > 1) {                                  // join: the last band of each render stream
      CU(cudaStreamWaitEvent(rs, scene->bandEvents[order[nBands - 1]], 0));
      CU(cudaStreamWaitEvent(rs, scene->bandEvents[order[nBands - 2]], 0));
    }
    CU(cudaEventRecord(scene->ev1, rs));
    unsigned long long c[3];
    CU(cudaMemcpyAsync(c, scene->d_counters, sizeof(c), cudaMemcpyDeviceToHost, rs));
    scene->bandProbes.assign(nBands, 0u);
    if (nBands) CU(cudaMemcpyAsync(scene->bandProbes.data(), scene->d_bandProbes, nBands * sizeof(uint32_t), cudaMemcpyDeviceToHost, rs));
    scene->bandKey = bandKey;
    CU(cudaStreamSynchronize(rs));
    CU(cudaStreamSynchronize(cp));
    if (timeline && nRows) {
      fprintf(stderr, "[yahr_b200 timeline] wall %.3f ms;", nowMs() - w0);
      for (uint32_t b = 0; b < nBands; ++b) {
        float tr = 0, tc = 0;
        cudaEventElapsedTime(&tr, scene->ev0, scene->bandEvents[b]);
        cudaEventElapsedTime(&tc, scene->ev0, copyDone[b]);
        fprintf(stderr, " band %u render %.3f copy %.3f;", b, tr, tc);
      }
      fprintf(stderr, "\n");
      for (auto e : copyDone) cudaEventDestroy(e);
    }
    {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, scene->ev0, scene->ev1) == cudaSuccess) scene->lastHostGpuMs = ms; else cudaGetLastError();
    }
    if (stats) {
      std::memset(stats, 0, sizeof(*stats));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, scene->ev0, scene->ev1));
      stats->n_primary = c[0]; stats->n_shadow = c[1]; stats->n_secondary = c[2];
      stats->gpu_ms = ms;
      if (plan.wavefront && ts.n) {
        for (int k = 0; k < 3; ++k) {
          float pm = 0;
          CU(cudaEventElapsedTime(&pm, scene->phaseEv[k], scene->phaseEv[k + 1]));
          stats->phase_ms[k] = pm;             // first band only
        }
      }
      stats->launches = launches;
      stats->tiles = ts.n;
      stats->h2d_bytes = (plan.wavefront ? sizeof(WavefrontParams) : sizeof(RenderParams)) * (uint64_t)launches;
      stats->d2h_bytes = d2h;
      stats->wall_ms = nowMs() - w0;
    }
    if (strategy) strategy->record(false, nowMs() - w0);
    return YAHR_OK;
  } catch (const CudaFailure& f) {
    return cudaFail(f);
  } catch (const std::exception& e) {
    return fail(YAHR_ERR_INTERNAL, e.what());
  }
}

int yahr_b200_render(yahr_scene* scene, const yahr_camera* cam, int recursion_depth, int spp, uint64_t seed,
                     float* rgb_out, uint32_t* primid_out, yahr_stats* stats) {
  if (!rgb_out) return fail(YAHR_ERR_INVALID_ARGUMENT, "rgb_out is NULL");
  return renderHost(scene, cam, recursion_depth, spp, seed, rgb_out, nullptr, primid_out, stats);
}

int yahr_b200_render_shard(yahr_scene* scene, const yahr_camera* cam, int recursion_depth, int spp, uint64_t seed,
                           int shard_index, int shard_count, float* rgb_out, uint32_t* primid_out, yahr_stats* stats) {
  if (!rgb_out) return fail(YAHR_ERR_INVALID_ARGUMENT, "rgb_out is NULL");
  return renderHost(scene, cam, recursion_depth, spp, seed, rgb_out, nullptr, primid_out, stats, shard_index, shard_count);
}

int yahr_b200_render_rgb8(yahr_scene* scene, const yahr_camera* cam, int recursion_depth, int spp, uint64_t seed,
                          unsigned char* rgb8_out, yahr_stats* stats) {
  if (!rgb8_out) return fail(YAHR_ERR_INVALID_ARGUMENT, "rgb8_out is NULL");
  return renderHost(scene, cam, recursion_depth, spp, seed, nullptr, rgb8_out, nullptr, stats);
}

int yahr_b200_render_shard_rgb8(yahr_scene* scene, const yahr_camera* cam, int recursion_depth, int spp, uint64_t seed,
                                int shard_index, int shard_count, unsigned char* rgb8_out, yahr_stats* stats) {
  if (!rgb8_out) return fail(YAHR_ERR_INVALID_ARGUMENT, "rgb8_out is NULL");
  return renderHost(scene, cam, recursion_depth, spp, seed, nullptr, rgb8_out, nullptr, stats, shard_index, shard_count);
}

// End-of-frame fence of the multi-GPU exchange without a collective: every pushing rank stores the frame's sequence
// number into ITS word of a flag array in rank 0's memory (mapped with yahr_b200_ipc_open) once its rows have landed;
// rank 0's stream waits until every word has reached the number.
int yahr_b200_flag_signal(uint32_t* d_flag, uint32_t value, void* stream) {
  if (!d_flag) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  try {
    CU(launchFlagSignal(d_flag, value, (cudaStream_t)stream));
    return YAHR_OK;
  } catch (const CudaFailure& f) {
    return cudaFail(f);
  }
}

int yahr_b200_flags_wait(uint32_t* d_flags, int first, int count, uint32_t value, void* stream) {
  if (!d_flags || first < 0 || count < 0 || count > 1024) return fail(YAHR_ERR_INVALID_ARGUMENT, "invalid argument");
  try {
    if (count) CU(launchFlagsWait(d_flags + first, count, value, (cudaStream_t)stream));
    return YAHR_OK;
  } catch (const CudaFailure& f) {
    return cudaFail(f);
  }
}

// Pinned host memory for the output buffers: device-to-host copies into it run at full PCIe speed
// and overlap with rendering (pageable memory works too, but the copies are then staged).
int yahr_b200_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(YAHR_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (yahr_b200_device_count() < 1) return fail(YAHR_ERR_NO_DEVICE, "no CUDA device available");
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(YAHR_ERR_OUT_OF_MEMORY, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return YAHR_OK;
}

int yahr_b200_host_free(void* p) {
  if (!p) return YAHR_OK;
  cudaError_t e = cudaFreeHost(p);
  if (e != cudaSuccess) return fail(YAHR_ERR_CUDA, std::string("cudaFreeHost: ") + cudaGetErrorString(e));
  return YAHR_OK;
}

int64_t yahr_b200_num_batches(int64_t num_threads, int64_t width, int64_t height) {
  return numBatches(num_threads, width, height);
}

int yahr_b200_batch_window(int64_t width, int64_t height, int64_t num, int64_t count, int32_t out[4]) {
  if (!out || count < 1 || num < 0 || num >= count || width < 1 || height < 1)
    return fail(YAHR_ERR_INVALID_ARGUMENT, "batch_window: invalid argument");
  TileWindow t = batchWindow(width, height, num, count);
  out[0] = t.x0; out[1] = t.y0; out[2] = t.x1; out[3] = t.y1;
  return YAHR_OK;
}

int yahr_b200_ipc_export(void* device_ptr, unsigned char handle_out[YAHR_B200_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == YAHR_B200_IPC_HANDLE_BYTES, "IPC handle size");
  if (!device_ptr || !handle_out) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, device_ptr);
  if (e != cudaSuccess) return fail(YAHR_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  std::memcpy(handle_out, &h, sizeof(h));
  return YAHR_OK;
}

int yahr_b200_ipc_open(const unsigned char handle[YAHR_B200_IPC_HANDLE_BYTES], void** device_ptr_out) {
  if (!handle || !device_ptr_out) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(device_ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(YAHR_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  return YAHR_OK;
}

int yahr_b200_ipc_close(void* device_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(device_ptr);
  if (e != cudaSuccess) return fail(YAHR_ERR_CUDA, std::string("cudaIpcCloseMemHandle: ") + cudaGetErrorString(e));
  return YAHR_OK;
}

// ---- .yahrr reader + expand (Scene.hs:15-86), output stage (main.hs:142) -------------------------
int yahr_b200_yahrr_load(const char* text, size_t length, yahr_loaded_scene** out) {
  if (!text || !out) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = nullptr;
  try {
    yahr_loaded_scene* h = new yahr_loaded_scene();
    std::string err;
    int rc = loadYahrr(std::string(text, length), h->scene, err);
    if (rc) { delete h; return fail(rc, err); }
    *out = h;
    return YAHR_OK;
  } catch (const std::exception& e) {
    return fail(YAHR_ERR_INTERNAL, e.what());
  }
}

void yahr_b200_yahrr_free(yahr_loaded_scene* h) { delete h; }

int yahr_b200_yahrr_describe(const yahr_loaded_scene* h, yahr_scene_desc* desc_out, yahr_camera* camera_out,
                             int* recursion_depth_out) {
  if (!h) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  if (desc_out) *desc_out = h->scene.desc();
  if (camera_out) *camera_out = h->scene.camera;
  if (recursion_depth_out) *recursion_depth_out = h->scene.recursionDepth;
  return YAHR_OK;
}

int yahr_b200_write_png_rgb8(const char* path, const unsigned char* rgb8, int width, int height) {
  if (!path || !rgb8 || width < 1 || height < 1) return fail(YAHR_ERR_INVALID_ARGUMENT, "invalid argument");
  std::string err;
  int rc = writePngRgb8(path, rgb8, width, height, err);
  return rc ? fail(rc, err) : YAHR_OK;
}

void yahr_b200_quantize_rgb8_host(const float* rgb, size_t count, unsigned char* out) {
  for (size_t i = 0; i < count; ++i) out[i] = quantize8(rgb[i]);
}

// ---- host-only inspection ----------------------------------------------------------------------
int yahr_b200_host_bvh_build(const yahr_scene_desc* desc, yahr_host_bvh** out) {
  if (!out) return fail(YAHR_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  try {
    std::vector<HostPrim> prims;
    std::vector<Box> bounds;
    std::string err;
    int rc = gatherPrimitives(desc, prims, bounds, err);
    if (rc) return fail(rc, err);
    yahr_host_bvh* h = new yahr_host_bvh();
    buildReferenceBvh(bounds, desc->bvh_max_depth, desc->split_mode, h->bvh);
    *out = h;
    return YAHR_OK;
  } catch (const std::exception& e) {
    return fail(YAHR_ERR_INTERNAL, e.what());
  }
}
void yahr_b200_host_bvh_destroy(yahr_host_bvh* b) { delete b; }
uint32_t yahr_b200_host_bvh_num_primitives(const yahr_host_bvh* b) { return b ? (uint32_t)b->bvh.order.size() : 0; }
uint32_t yahr_b200_host_bvh_num_nodes(const yahr_host_bvh* b) { return b ? (uint32_t)b->bvh.nodes.size() : 0; }
uint32_t yahr_b200_host_bvh_depth(const yahr_host_bvh* b) { return b ? b->bvh.depth : 0; }
int yahr_b200_host_bvh_order(const yahr_host_bvh* b, uint32_t* order_out) {
  if (!b || !order_out) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  std::memcpy(order_out, b->bvh.order.data(), b->bvh.order.size() * sizeof(uint32_t));
  return YAHR_OK;
}
int yahr_b200_host_bvh_preorder(const yahr_host_bvh* b, uint32_t* kinds, uint32_t* firsts, uint32_t* counts,
                                float* boxes6) {
  if (!b || !kinds || !firsts || !counts || !boxes6) return fail(YAHR_ERR_INVALID_ARGUMENT, "NULL argument");
  for (size_t i = 0; i < b->bvh.nodes.size(); ++i) {
    const RefNode& n = b->bvh.nodes[i];
    kinds[i] = n.kind; firsts[i] = n.first; counts[i] = n.kind == kInner ? 0 : n.count;
    float* bx = boxes6 + 6 * i;
    bx[0] = n.box.lo.x; bx[1] = n.box.lo.y; bx[2] = n.box.lo.z; bx[3] = n.box.hi.x; bx[4] = n.box.hi.y; bx[5] = n.box.hi.z;
  }
  return YAHR_OK;
}
int yahr_b200_camera_matrices(const yahr_camera* cam, float tf16[16], float vtf16[16]) {
  CameraSetup cs;
  std::string err;
  int rc = setupCamera(cam, cs, err);
  if (rc) return fail(rc, err);
  std::memcpy(tf16, cs.tf.m, 64);
  std::memcpy(vtf16, cs.vtf.m, 64);
  return YAHR_OK;
}

}  // extern "C"
