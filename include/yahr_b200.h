/* yahr_b200.h -- C ABI of libyahr_b200.so, the B200 (sm_100a) implementation of yahr's per-pixel
 * render loop.
 *
 * The reference (pstiasny/yahr, pure Haskell) has NO FFI or plugin API.  The only operator-shaped
 * seam on this path is the function type shared by its three renderers
 *     render / renderEval / renderPar ::
 *         Int -> Int -> [[SampleCoordinates]] -> (Float -> Float -> Ray) -> (Ray -> Spectrum) -> [Sample]
 * (main.hs:68-70, 75-77, 86-88), selected by the `-p/--parallel-mode` string (main.hs:38, 133-137)
 * and consumed by samplesToImage (main.hs:98-107).  Closures cannot cross a C ABI, so a GPU mode
 * takes the DATA those closures close over: camera, integrator, cullingMode, lights, materials and
 * `objects >>= expand` (main.hs:41-60, 117-126).  That is what the entry points below accept.
 * INTEGRATION.md shows the `foreign import ccall` binding and the `"gpu" ->` case a maintainer
 * would add at main.hs:134-137.
 *
 * Conventions
 *  - plain C, no C++ types; every function returns 0 on success and a non-zero yahr_status on
 *    error (message via yahr_b200_last_error(), thread-local).  Nothing throws or exits.
 *  - the caller owns every input array (copied during scene_create; may be freed afterwards) and
 *    every output buffer.  The library owns device memory behind the opaque handle.
 *  - a handle is used by one host thread at a time; distinct handles are independent.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *  - all floating point is IEEE binary32 like the reference's `Float` (Vectors.hs:5); the kernels
 *    are compiled without FMA contraction so that primary-hit primitive IDs are bit-exact.
 */
#ifndef YAHR_B200_H
#define YAHR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YAHR_B200_ABI_VERSION 2

typedef enum yahr_status {
  YAHR_OK = 0,
  YAHR_ERR_INVALID_ARGUMENT = 1,
  YAHR_ERR_NO_DEVICE = 2,        /* no CUDA device / driver: there is no CPU fallback */
  YAHR_ERR_CUDA = 3,
  YAHR_ERR_OUT_OF_MEMORY = 4,
  YAHR_ERR_UNKNOWN_MATERIAL = 5, /* material index out of range (the reference: Map.! error, main.hs:55) */
  YAHR_ERR_BVH_TOO_DEEP = 6,     /* built tree deeper than the traversal stack (YAHR_B200_MAX_STACK) */
  YAHR_ERR_NON_FINITE_INPUT = 7, /* NaN/inf in geometry: behaviour of the reference is unspecified */
  YAHR_ERR_INTERNAL = 8,
  YAHR_ERR_PARSE = 9,            /* .yahrr text not accepted (the reference: "Prelude.read: no parse") */
  YAHR_ERR_IO = 10
} yahr_status;

#define YAHR_B200_MAX_STACK 64
#define YAHR_B200_MAX_RECURSION 16
#define YAHR_PRIM_MISS 0xFFFFFFFFu

/* CullingMode = BVH Int SplitMode (Culling.hs:18-19) */
enum { YAHR_SPLIT_MIDPOINT = 0, YAHR_SPLIT_SAH = 1 };

/* The scene as data: what buildCollisionModel (main.hs:41-60) and `radiance` (main.hs:125-126)
 * close over.  Primitives are ALREADY EXPANDED (Scene.expand, Scene.hs:61-86) and listed in
 * `objects >>= expand` order (main.hs:44); the position in that order is the PRIMITIVE ID
 * reported by the render calls. */
typedef struct yahr_scene_desc {
  uint32_t n_triangles;
  const float* tri_p0;           /* n_triangles x 3  (Scene.Triangle p0, Scene.hs:20-28) */
  const float* tri_p1;
  const float* tri_p2;
  const float* tri_n0;           /* vertex shading normals, NOT normalised by the library */
  const float* tri_n1;
  const float* tri_n2;
  const uint32_t* tri_material;  /* n_triangles, index into materials */
  uint32_t n_spheres;
  const float* sph_center;       /* n_spheres x 3    (Scene.Sphere position, Scene.hs:16-19) */
  const float* sph_radius;       /* n_spheres */
  const uint32_t* sph_material;
  /* Order of the mixed primitive list: NULL = all spheres, then all triangles; otherwise
   * n_spheres + n_triangles entries, entry i = (kind << 31) | index, kind 0 sphere / 1 triangle. */
  const uint32_t* prim_order;
  uint32_t n_materials;
  const float* materials;        /* n_materials x 7: diffuse rgb, specular rgb, shininess
                                    (Scene.BlinnPhongMaterial, Scene.hs:45-50; `ambient` is ignored by
                                    Shaders.blinnPhong, Shaders.hs:12-14) */
  uint32_t n_lights;
  const float* lights;           /* n_lights x 6: position xyz, spectrum rgb (Lights.PointLight, Lights.hs:7) */
  int32_t bvh_max_depth;         /* BVH <maxDepth> _   (Culling.hs:19) */
  int32_t split_mode;            /* YAHR_SPLIT_*       (Culling.hs:18) */
  /* EXTENSION -- no reference counterpart (the reference has point lights only, Lights.hs:7): one-sided
   * parallelogram lights.  n_area_lights x 13 floats: corner xyz, edge1 xyz, edge2 xyz, radiance rgb,
   * samples (>= 1, as a float).  Every sample is a virtual point light of spectrum
   * cos_l * (area / samples) * radiance at a counter-based pseudo-random point of the parallelogram, fed
   * through the reference's own direct-lighting arithmetic (csrc/render_device.cuh lightSampleU).
   * 0 / NULL = none. */
  uint32_t n_area_lights;
  const float* area_lights;
} yahr_scene_desc;

/* Cameras.Camera (Cameras.hs:54-56).  Image size is floor(imW) x floor(imH) (main.hs:122-123). */
typedef struct yahr_camera {
  float imW, imH, focalLength;
  float lookDir[3], upDir[3], position[3];
} yahr_camera;

/* Traversal order.  REFERENCE walks the tree exactly like the closure tree of Culling.hs:24-25
 * (left child first, box test on entry with the current tMax): primitive IDs are bit-exact by
 * construction.  ORDERED visits the nearer child first and resolves equal-t ties by DFS rank
 * (SURVEY.md note T); it can differ from the reference only where a float slab test is "inverted"
 * with respect to the hit it bounds (documented exact-tie class). */
enum { YAHR_TRAVERSAL_REFERENCE = 0, YAHR_TRAVERSAL_ORDERED = 1 };

typedef struct yahr_render_opts {
  int32_t recursion_depth;  /* WhittedIntegrator { recursionDepth } (Integrators.hs:18-20); 1 = direct lighting */
  int32_t spp;              /* samples per pixel; sample 0 is the reference's ray through the integer
                               raster coordinate (main.hs:73); samples >= 1 are an extension */
  uint64_t seed;            /* seed of the counter-based jitter for samples >= 1 */
  int32_t traversal;        /* YAHR_TRAVERSAL_* */
  int32_t tile_stride;      /* render tiles tile_offset, tile_offset + tile_stride, ... of the reference's */
  int32_t tile_offset;      /*   own squareBatches tiling (Sampling.hs:5-21); 1 / 0 = the whole image      */
  int32_t kernel;           /* 0 = default kernel set; 1 = megakernel, 2 = wavefront kernels (experiments) */
  int32_t reserved[4];      /* 0 = defaults.  [0]: tuning word of the wavefront kernels (experiments; DESIGN.md
                               section 10), [1]: 1 = tile_stride / tile_offset count whole ROWS of the tile grid */
} yahr_render_opts;

typedef struct yahr_stats {
  uint64_t n_primary;       /* camera rays traced */
  uint64_t n_shadow;        /* shadow probes traced (only where lensq k > 0, Integrators.hs:59) */
  uint64_t n_secondary;     /* reflection rays traced (recursion_depth > 1) */
  double gpu_ms;            /* CUDA-event time of the kernels of this call */
  double wall_ms;           /* host wall time of the call (copies included for the host-buffer entry) */
  uint64_t h2d_bytes, d2h_bytes;
  uint32_t launches;        /* kernel launches issued by this call */
  uint32_t tiles;           /* reference tiles rendered by this call */
  double phase_ms[4];       /* wavefront set: primary trace, shade, shadow trace (first sample pass);
                               megakernel: [0] = the kernel */
} yahr_stats;

typedef struct yahr_scene_info {
  uint32_t n_primitives, n_nodes, n_multi_leaves, depth;
  uint64_t device_bytes;    /* HBM held by the scene */
  double build_ms;          /* host BVH construction (restating Culling.hs:27-112) */
  double upload_ms;
  uint32_t built_on_device; /* 1: BVH built by the GPU builder, 0: host builder */
  uint32_t n_wide_nodes;    /* 4-wide collapse of the tree used by the traversal kernels (0: not built) */
} yahr_scene_info;

typedef struct yahr_scene yahr_scene;

/* --- library ---------------------------------------------------------------------------------- */
int yahr_b200_abi_version(void);
/* sizeof the ABI structs as the library was compiled (0 yahr_scene_desc, 1 yahr_camera, 2 yahr_render_opts, 3 yahr_stats,
 * 4 yahr_scene_info), so that a foreign binding can check its own marshalling; -1 for an unknown index. */
int yahr_b200_sizeof(int which);
int yahr_b200_device_count(void);                 /* 0 when no CUDA device is usable */
const char* yahr_b200_last_error(void);           /* thread-local, never NULL */

/* --- scene: replaces buildCollisionModel + cull (main.hs:41-60, 118; Culling.hs:21-112) -------- */
/* Builds the reference's BVH (same split decisions, same topology, same left-first DFS primitive order, both split
 * modes) on the CURRENT CUDA device from the uploaded primitive arrays, collapses it into the 4-wide tree the
 * traversal kernels walk, and keeps everything resident.  Degenerate inputs the device builder declines, and
 * YAHR_B200_HOST_BUILD=1, take the host twin of the same algorithm. */
int yahr_b200_scene_create(const yahr_scene_desc* desc, yahr_scene** out);
void yahr_b200_scene_destroy(yahr_scene* scene);
int yahr_b200_scene_info(const yahr_scene* scene, yahr_scene_info* out);
/* Inspection: download the device-resident BVH.  order_out: n_primitives uint32 (primitive ID per DFS
 * position); nodes_out: n_nodes x 16 floats (64-byte traversal nodes); multi_out: n_multi_leaves x 2
 * uint32 (first, count).  Any output may be NULL. */
int yahr_b200_scene_download_bvh(const yahr_scene* scene, uint32_t* order_out, float* nodes_out, uint32_t* multi_out,
                                 uint32_t* root_ref_out, float root_box_out[6]);
/* Inspection: the 4-wide collapse of that tree the traversal kernels walk (yahr_scene_info.n_wide_nodes x 32
 * floats; layout in csrc/device_types.cuh). */
int yahr_b200_scene_download_wide(const yahr_scene* scene, float* wide_out);

/* Inspection: the COMPRESSED copy of the wide nodes (n_wide_nodes x 16 words: origin' xyz + step exponents, 24 grid
 * bytes + 8 spare, four child refs; csrc/wide_bvh.cu) and the exact boxes the compressed walk tests at the leaves
 * (n_primitives x 8 floats by DFS position, n_multi_leaves x 8 floats: lo.xyz hi.x | hi.yz - -).  Any pointer may be
 * NULL.  Error when the scene has no compressed nodes (non-finite box coordinates, or no wide tree). */
int yahr_b200_scene_download_compressed(const yahr_scene* scene, float* cwide_out, float* leaf_box_out,
                                        float* multi_box_out);

/* --- render: replaces render / renderEval / renderPar + samplesToImage (main.hs:68-107) -------- */
/* Host-buffer entry (the call the Haskell host makes).  rgb_out: W*H*3 floats, row-major, RGB
 * interleaved, row 0 = top -- exactly JuicyPixels' `Image PixelRGBF` storage (main.hs:98-107).
 * primid_out: W*H uint32 primary-hit primitive IDs (YAHR_PRIM_MISS for background) or NULL.
 * Whitted integrator with `recursion_depth`; spp samples per pixel. */
int yahr_b200_render(yahr_scene* scene, const yahr_camera* cam, int recursion_depth, int spp, uint64_t seed,
                     float* rgb_out, uint32_t* primid_out, yahr_stats* stats);

/* Device-buffer entry: same computation, outputs stay in device memory (d_rgb: W*H*3 floats,
 * d_primid: W*H uint32 or NULL; either may point into a PEER GPU's memory mapped with
 * yahr_b200_ipc_open -- pixels are then stored over NVLink as they are produced).  Only the pixels
 * of the selected tiles are written.  `stream` is a cudaStream_t (NULL = default stream).  The call
 * returns after enqueueing unless `stats` is non-NULL (then it synchronises the stream). */
int yahr_b200_render_device(yahr_scene* scene, const yahr_camera* cam, const yahr_render_opts* opts,
                            float* d_rgb, uint32_t* d_primid, void* stream, yahr_stats* stats);

/* Measurement aid (no reference counterpart; not on the product path): the same frame as yahr_b200_render_device
 * through the COUNTING build of the wavefront kernels -- identical results plus the GPU's own work counters, from
 * which bench.py derives the bytes the kernels really request per ray (SURVEY.md 8d counts them under the reference's
 * traversal, main.hs:73 -> Culling.hs:24-25; the 4-wide walk skips ancestor boxes).  recursion_depth 1 only.
 * counts_out[0..7] closest-hit walks, [8..15] any-hit walks, each {4-wide node visits (128 B), binary node visits
 * (64 B), primitive tests (48 B), normal fetches (48 B), stack pushes (8 B), stack pops (8 B), -, shaded hits}. */
int yahr_b200_render_device_counted(yahr_scene* scene, const yahr_camera* cam, const yahr_render_opts* opts,
                                    float* d_rgb, uint32_t* d_primid, void* stream, yahr_stats* stats,
                                    uint64_t counts_out[16]);

/* Multi-GPU device-buffer entry.  Renders the rows shard_index, shard_index + shard_count, ... of the reference's
 * tile grid into the LOCAL buffers (full-frame sized, on the scene's GPU), then pushes exactly those pixel rows into
 * the gather buffers with device-to-device copies enqueued on `stream` (the gather frame usually lives on rank 0 and
 * is mapped here with yahr_b200_ipc_open: a few large NVLink transfers per frame).  gather == local or NULL: no push.
 * opts->tile_stride / tile_offset are ignored.  Returns after enqueueing unless `stats` is non-NULL. */
int yahr_b200_render_device_shard(yahr_scene* scene, const yahr_camera* cam, const yahr_render_opts* opts,
                                  int shard_index, int shard_count, float* d_rgb_local, float* d_rgb_gather,
                                  uint32_t* d_primid_local, uint32_t* d_primid_gather, void* stream, yahr_stats* stats);

/* Multi-GPU host-buffer entry.  The frame is cut into whole rows of the reference's tile grid (squareBatches,
 * Sampling.hs:5-21), dealt to the shards in blocks of up to four consecutive rows (block b -> shard b mod
 * shard_count; the shards of one frame partition it for every shard_count); this call renders the blocks of
 * shard_index on the scene's GPU and copies exactly those pixel rows into rgb_out / primid_out, which are FULL-frame
 * buffers (W*H*3 floats, W*H uint32), typically one pinned allocation shared by all shards.  The host runs one call
 * per GPU concurrently (one thread or process each, each with its own yahr_scene on its own device), so every GPU
 * uses its own PCIe link and no inter-GPU exchange is needed.  shard_count = 1 is yahr_b200_render. */
int yahr_b200_render_shard(yahr_scene* scene, const yahr_camera* cam, int recursion_depth, int spp, uint64_t seed,
                           int shard_index, int shard_count, float* rgb_out, uint32_t* primid_out, yahr_stats* stats);

/* The same with the reference's 8-bit output stage (main.hs:142) applied on the GPU: rgb8_out is the FULL-frame
 * W*H*3 byte buffer shared by the shards; a quarter of the bytes cross PCIe. */
int yahr_b200_render_shard_rgb8(yahr_scene* scene, const yahr_camera* cam, int recursion_depth, int spp, uint64_t seed,
                                int shard_index, int shard_count, unsigned char* rgb8_out, yahr_stats* stats);

/* End-of-frame fence of the multi-GPU device-buffer exchange (the reference's `concat`, main.hs:83,95, has no
 * counterpart for it: its sparks share one heap).  A pushing rank enqueues yahr_b200_flag_signal on the stream its
 * stores / copies into rank 0's frame ran on: `value` (the frame's sequence number, growing) is written to d_flag --
 * this rank's word of an array in rank 0's memory, mapped with yahr_b200_ipc_open -- after a system-scope fence.
 * Rank 0 enqueues yahr_b200_flags_wait on the stream that consumes the frame: the stream blocks until the `count`
 * words from `first` on have all reached `value` (it gives up after ~4 s rather than hang the device). */
int yahr_b200_flag_signal(uint32_t* d_flag, uint32_t value, void* stream);
int yahr_b200_flags_wait(uint32_t* d_flags, int first, int count, uint32_t value, void* stream);

/* Same frame with the reference's output stage applied on the GPU: JuicyPixels' ImageRGBF -> 8-bit
 * conversion used by savePngImage (main.hs:142), truncate (255 * max 0 (min 1 x)), no gamma.
 * rgb8_out: W*H*3 bytes, row-major, row 0 = top.  Only a quarter of the bytes cross PCIe. */
int yahr_b200_render_rgb8(yahr_scene* scene, const yahr_camera* cam, int recursion_depth, int spp, uint64_t seed,
                          unsigned char* rgb8_out, yahr_stats* stats);

/* --- scene file: replaces `read sceneFile :: S.Scene` + `S.objects s >>= S.expand` ----------------
 * (main.hs:115-117, 44; Scene.hs:15-86).  Host-only, no CUDA needed.  `text` is the derived-Read
 * syntax of Scene (the .yahrr format).  YAHR_ERR_PARSE mirrors "Prelude.read: no parse",
 * YAHR_ERR_UNKNOWN_MATERIAL mirrors the Map.! failure of main.hs:55. */
typedef struct yahr_loaded_scene yahr_loaded_scene;
int yahr_b200_yahrr_load(const char* text, size_t length, yahr_loaded_scene** out);
void yahr_b200_yahrr_free(yahr_loaded_scene* loaded);
/* Fills a descriptor that points INTO `loaded` (valid until yahr_b200_yahrr_free), the camera and the
 * integrator's recursionDepth; any output may be NULL. */
int yahr_b200_yahrr_describe(const yahr_loaded_scene* loaded, yahr_scene_desc* desc_out, yahr_camera* camera_out,
                             int* recursion_depth_out);

/* --- output stage helpers (main.hs:142) ---------------------------------------------------------- */
int yahr_b200_write_png_rgb8(const char* path, const unsigned char* rgb8, int width, int height);
void yahr_b200_quantize_rgb8_host(const float* rgb, size_t count, unsigned char* out);

/* Pinned host memory for output buffers (optional): device-to-host copies into it run at full PCIe
 * speed and overlap with rendering inside yahr_b200_render.  Any host pointer works as an output. */
int yahr_b200_host_alloc(size_t bytes, void** out);
int yahr_b200_host_free(void* p);

/* --- tile arithmetic of the reference (Sampling.hs:9-21, main.hs:109-110, 128-131) ------------- */
int64_t yahr_b200_num_batches(int64_t num_threads, int64_t width, int64_t height);
int yahr_b200_batch_window(int64_t width, int64_t height, int64_t num, int64_t count, int32_t out_x0y0x1y1[4]);

/* --- peer-memory plumbing for the one-process-per-GPU driver ----------------------------------- */
#define YAHR_B200_IPC_HANDLE_BYTES 64
int yahr_b200_ipc_export(void* device_ptr, unsigned char handle_out[YAHR_B200_IPC_HANDLE_BYTES]);
int yahr_b200_ipc_open(const unsigned char handle[YAHR_B200_IPC_HANDLE_BYTES], void** device_ptr_out);
int yahr_b200_ipc_close(void* device_ptr);

/* --- host-only inspection (no CUDA needed): the BVH the library would upload ------------------- */
typedef struct yahr_host_bvh yahr_host_bvh;
int yahr_b200_host_bvh_build(const yahr_scene_desc* desc, yahr_host_bvh** out);
void yahr_b200_host_bvh_destroy(yahr_host_bvh* bvh);
uint32_t yahr_b200_host_bvh_num_primitives(const yahr_host_bvh* bvh);
uint32_t yahr_b200_host_bvh_num_nodes(const yahr_host_bvh* bvh);      /* reference-tree nodes, pre-order */
uint32_t yahr_b200_host_bvh_depth(const yahr_host_bvh* bvh);
/* order_out[i] = primitive ID at left-first DFS leaf position i */
int yahr_b200_host_bvh_order(const yahr_host_bvh* bvh, uint32_t* order_out);
/* pre-order dump of the reference tree: kind (0 null, 1 single leaf, 2 inner, 3 multi-leaf),
 * first/count (leaf range in DFS order) and the node's box (6 floats) */
int yahr_b200_host_bvh_preorder(const yahr_host_bvh* bvh, uint32_t* kinds, uint32_t* firsts, uint32_t* counts,
                                float* boxes6);
/* the two 4x4 matrices of Cameras.computeInitialRay (Cameras.hs:80-81), row-major */
int yahr_b200_camera_matrices(const yahr_camera* cam, float tf16[16], float vtf16[16]);

#ifdef __cplusplus
}
#endif
#endif /* YAHR_B200_H */
