/* rt_b200.h -- C ABI of the B200-native render hot path (librt_b200.so).
 *
 * This is the drop-in boundary for ONE path of nikitakaraevv/ray-tracing-engine: everything
 * `Renderer::render(Image&)` does per pixel sample (reference: source/Renderer.h:36,
 * source/Renderer.cpp:203-272, called once from source/Main.cpp:224) plus the two Renderer
 * constructors that parameterise it (source/Renderer.h:19,23; source/Renderer.cpp:15-31).
 * The reference has no FFI of its own (it is one C++ translation unit); the seam is introduced at
 * that call.  INTEGRATION.md shows the few lines a maintainer adds to Main.cpp to route
 * `renderer.render(image)` through rt_render().
 *
 * Conventions
 *   - plain C, POD structs, caller-owned host buffers unless a name ends in _device;
 *   - every function returns RT_OK (0) or a negative rt_status; rt_last_error() gives the text of the
 *     calling thread's last failure (the reference throws std::runtime_error / std::logic_error or
 *     calls exit(1); the CLI maps non-zero statuses back to the same messages + exit(1));
 *   - no CPU fallback: when no CUDA device is usable every entry point fails with RT_ERR_NO_DEVICE;
 *   - one host thread per context, blocking calls; one context per GPU (one process per GPU).
 */
#ifndef RT_B200_H
#define RT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rt_status {
  RT_OK = 0,
  RT_ERR_INVALID = -1,      /* bad argument (null pointer, negative count, k > RT_MAX_K, ...)         */
  RT_ERR_NO_DEVICE = -2,    /* no usable CUDA device / driver                                         */
  RT_ERR_CUDA = -3,         /* a CUDA runtime call or kernel failed; see rt_last_error()              */
  RT_ERR_EMPTY_TREE = -4,   /* kdtree.h:181  "tree is empty"                                          */
  RT_ERR_K_TOO_LARGE = -5,  /* kdtree.h:182-183 "k is greater than the number of nodes"               */
  RT_ERR_OOM = -6
} rt_status;

/* Material{kd, alpha, albedo, F0}: source/Material.h:8-20,62-64 */
typedef struct rt_material {
  float kd, alpha;
  float albedo[3];
  float f0[3];
} rt_material;

/* LightSource: source/LightSource.h:19-33,61-65.  normal/vertical/horizontal are the HOST-computed
 * basis (LightSource.h:29-32); the device never recomputes them. */
typedef struct rt_light {
  float position[3], color[3], normal[3], vertical[3], horizontal[3];
  float intensity, side, ac, al, aq, factor;
} rt_light;

/* Camera: source/Camera.h:9-24,34-40 (position, lower-left corner, horizontal, vertical). */
typedef struct rt_camera {
  float position[3], lower_left[3], horizontal[3], vertical[3];
} rt_camera;

/* Scene::meshes()/lightsources()/camera() flattened (source/Scene.h:14-31, source/Mesh.h:136-139).
 * Mesh order and triangle order are preserved: they are the reference's tie-break order
 * (source/RayTracer.h:32-51). */
typedef struct rt_scene {
  int32_t num_vertices, num_triangles, num_meshes, num_lights;
  const float* positions;              /* 3*V                                                   */
  const float* normals;                /* 3*V                                                   */
  const int32_t* triangles;            /* 3*T, GLOBAL vertex indices                            */
  const int32_t* mesh_first_triangle;  /* M+1 offsets into triangles                            */
  const int32_t* mesh_first_vertex;    /* M+1 offsets into positions (to recover mesh-local ids) */
  const rt_material* materials;        /* M                                                     */
  const rt_light* lights;              /* L                                                     */
  rt_camera camera;
} rt_scene;

/* Renderer constructor arguments (source/Renderer.cpp:15-31) + image size + our additive knobs. */
typedef struct rt_params {
  int32_t width, height;
  int32_t num_rays;     /* -N: samples per pixel (m_numRays)                                     */
  int32_t mode;         /* -m: 0 ray trace, 1 path trace (depth 3, Renderer.cpp:246,249)         */
  int32_t num_photons;  /* -p: requested photons, 0 = no photon map                              */
  int32_t k;            /* -k: neighbours gathered (<= RT_MAX_K)                                 */
  uint64_t seed;        /* random stream seed (DESIGN.md "RNG contract")                         */
  /* sharding of the pixel grid across ranks: tile (tx,ty) of shard_tile x shard_tile pixels belongs
   * to rank (ty*tiles_x + tx) % shard_count.  shard_count <= 1 renders everything. */
  int32_t shard_rank, shard_count, shard_tile;
  /* sharding by sample index: render samples [sample_first, sample_first + sample_count) of the
   * num_rays-sample frame (sample_count <= 0: all of them).  The stratum of a sample depends on its
   * GLOBAL index (source/RayTracer.h:111-115), so num_rays stays the global N on every rank.
   * sample_first == num_rays with sample_count == 0 is the empty range (a rank with no share). */
  int32_t sample_first, sample_count;
  /* samples traced per wavefront batch; 0 = choose from free HBM */
  int32_t samples_per_batch;
  /* BVH box padding as a fraction of the scene extent; 0 = default (2^-14) */
  float bvh_pad;
  int32_t flags;        /* RT_FLAG_* */
} rt_params;

/* kdtree::knearest accepts any k <= number of nodes (source/kdtree.h:180-183) and Scene::lightsources() any number
 * of lights (source/Scene.h:14-26, source/Renderer.cpp:49); the limits below only bound index arithmetic.  Up to
 * k = 64 the candidates of a query live in shared memory, beyond that in a global scratch (slower, same results). */
#define RT_MAX_K 4096
#define RT_MAX_LIGHTS 4096
#define RT_FLAG_BRUTE_FORCE 1 /* trace with the O(T) scan instead of the BVH (parity hook) */
#define RT_FLAG_KNN_EXACT 2   /* gather the canonical exact k nearest photons (by distance, then array index) instead of
                                 reproducing kdtree::knearest's two quirks: fewer node visits, but 0.5-2.8 % of the
                                 queries -- and the pixels they feed -- differ from the reference */

/* one ray / one nearest-hit record, as the parity hooks exchange them */
typedef struct rt_ray {
  float origin[3];
  float direction[3];
} rt_ray;
typedef struct rt_hit {
  int32_t triangle; /* global triangle index in scene order, -1 = miss                          */
  float u, v, t;    /* the reference's u, v, d (source/RayTracer.h:45-47)                        */
} rt_hit;

/* Particle: source/Particle.h:33-35 (28 bytes) */
typedef struct rt_photon {
  float position[3];
  float direction[3]; /* incomeDirection */
  float weight;
} rt_photon;

enum {
  RT_KERNEL_RAYGEN = 0,        /* k_raygen                                                      */
  RT_KERNEL_TRACE_NEAREST = 1, /* k_trace_sp8u3<nearest>                                        */
  RT_KERNEL_SORT = 2,          /* k_sort_count / k_sort_scan / k_sort_scatter                   */
  RT_KERNEL_SHADE = 3,         /* k_shade (direct lighting / photon shading)                    */
  RT_KERNEL_TRACE_ANY = 4,     /* k_trace_sp8u3<any-hit> over the shadow rays                   */
  RT_KERNEL_COMBINE = 5,       /* k_combine                                                     */
  RT_KERNEL_RESOLVE = 6,       /* k_resolve                                                     */
  RT_KERNEL_EMIT = 7,          /* k_emit (photon emission)                                      */
  RT_KERNEL_OTHER = 8,         /* scatter / composite                                           */
  RT_KERNEL_GATHER = 9,        /* k_knn_gather (persistent k-nearest-photon gather)             */
  RT_NUM_KERNEL_CLASSES = 10
};

typedef struct rt_stats {
  uint64_t rays;           /* logical RayTracer::rayTrace invocations (primary+bounce+shadow+photon) */
  uint64_t primary_rays, bounce_rays, shadow_rays, photon_rays;
  uint64_t knn_queries;
  uint64_t samples;        /* pixel samples rendered                                                */
  uint64_t kernel_launches;/* launches of OUR kernels since rt_create / rt_reset_stats               */
  double device_ms;        /* CUDA-event time of the kernels of the last rt_render* call             */
  double trace_ms;         /* ... of which the trace/shade wavefront kernels                         */
  double photon_ms;        /* CUDA-event time of the last photon emission                            */
  int32_t bvh_nodes, bvh_depth;
  int64_t photons_stored;
  double create_ms;        /* host wall time of rt_create ...                                        */
  double bvh_build_ms;     /* ... of which the host BVH build                                        */
  double kd_build_ms;      /* host wall time of the last kd-tree build (rt_set_photons)              */
  uint64_t kd_visits;      /* kd-tree nodes visited by the k-NN queries: kdtree::visited() summed, minus
                              the visits the exact plane-distance bound skips                        */
  /* CUDA-event time (ms) and launch count of every kernel class (RT_KERNEL_*), summed over the rt_render* calls since
   * rt_create / rt_reset_stats; the events sit on the launching stream around each launch.  device_ms_total is
   * device_ms summed the same way. */
  double kernel_ms[RT_NUM_KERNEL_CLASSES];
  uint64_t kernel_count[RT_NUM_KERNEL_CLASSES];
  double device_ms_total;
} rt_stats;

typedef struct rt_ctx rt_ctx;

const char* rt_last_error(void);
int rt_device_count(void);
/* sizeof of {rt_material, rt_light, rt_camera, rt_scene, rt_params, rt_ray, rt_hit, rt_photon, rt_stats} as this
 * library was compiled (pure host function): lets a foreign-language binding verify its struct mirrors. */
int rt_abi_sizes(int32_t* out, int32_t capacity);

/* Copies the scene to the device (caller keeps ownership of every host array) and builds the BVH with
 * the reference's split policy (source/BVH.h:100-161): on the device from 8192 triangles (csrc/bvh_build.cu), on
 * the host below that (csrc/host_build.cpp); both produce the same tree up to the order of equal keys.
 * `device` is the CUDA ordinal (one context per GPU). */
int rt_create(const rt_scene* scene, const rt_params* params, int device, rt_ctx** out);
int rt_destroy(rt_ctx* ctx);
int rt_set_params(rt_ctx* ctx, const rt_params* params); /* new size / N / mode / k / seed / shard */

/* rt_composite on the device: sum_rgb_device / counter_device are full-frame device buffers (e.g. the result of
 * the NCCL reduce of every rank's rt_render_accumulate_device), rgb_inout a HOST buffer holding the background on
 * entry and the composite on exit.  Same arithmetic as rt_composite, bit for bit.  The kernel runs on the context's
 * own stream: the caller must have synchronised with whatever produced the device buffers (e.g. the NCCL stream). */
int rt_composite_device(rt_ctx* ctx, int32_t num_rays, const float* sum_rgb_device, const int32_t* counter_device,
                        float* rgb_inout);

/* Replaces `renderer.render(image)` (source/Main.cpp:224; source/Renderer.cpp:203-272).
 * rgb_inout: W*H*3 floats, row-major, y = 0 is the top row (source/Image.h:23-29).  In: the
 * background (Image::fillBackground, source/Image.cpp:12-21).  Out: the final composite
 * (source/Renderer.cpp:262-265,271).  Builds the photon map first when num_photons > 0
 * (source/Renderer.cpp:209-213).  With shard_count > 1 only the owned pixels are composited; the
 * rest are returned unchanged. */
int rt_render(rt_ctx* ctx, float* rgb_inout);

/* rt_render with the reference's progressive preview (source/Renderer.cpp:262-269: after pass i the image
 * `updateImage/(i+1) + background*(i+1-counter)/(i+1)` is written to update.ppm).  `fn` is called after every
 * `every` sample passes (and after the last one) with that composite over the whole frame; the snapshot buffer
 * belongs to the library and is valid only during the call.  The samples are accumulated in index order whatever
 * `every` is, so the final image is bit-identical to rt_render's.  every <= 0 or fn == NULL: plain rt_render. */
typedef void (*rt_progress_fn)(void* user, int32_t samples_done, int32_t num_rays, const float* rgb_snapshot);
int rt_render_progressive(rt_ctx* ctx, float* rgb_inout, int32_t every, rt_progress_fn fn, void* user);

/* The same work without the composite: per-pixel sums of the clamped sample colours
 * (`updateImage`, source/Renderer.cpp:254-258) and hit counters (`counter`, :255-257), W*H*3 floats
 * and W*H int32, row-major.  Pixels owned by other shards are written as zero so that a sum-reduce
 * over ranks assembles the frame.  The _device variant takes device pointers on ctx's GPU (e.g.
 * torch tensors) and leaves the result there for an NCCL reduce. */
int rt_render_accumulate(rt_ctx* ctx, float* sum_rgb, int32_t* counter);
int rt_render_accumulate_device(rt_ctx* ctx, float* sum_rgb_device, int32_t* counter_device);
/* The same as ONE buffer of W*H float4 {sum_r, sum_g, sum_b, counter} (device pointer, 16-byte aligned), so that a
 * multi-GPU frame needs a single sum-reduce: the counter is exact as a binary32 for num_rays < 2^24.
 * rt_composite_packed_device is rt_composite_device on that layout. */
int rt_render_accumulate_packed_device(rt_ctx* ctx, float* sum_rgbn_device);
int rt_composite_packed_device(rt_ctx* ctx, int32_t num_rays, const float* sum_rgbn_device, float* rgb_inout);
/* `saveImage = update/N + background*(N-counter)/N` (source/Renderer.cpp:262-265) on the host. */
int rt_composite(int32_t width, int32_t height, int32_t num_rays, const float* sum_rgb, const int32_t* counter,
                 float* rgb_inout);

/* Per-sample outputs for parity tests: clamped colour (3 floats) and posIntersectionFound (1 byte)
 * of samples [s0,s1) over the window [x0,x1)x[y0,y1), laid out [sample][y][x]. */
int rt_render_samples(rt_ctx* ctx, int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t s0, int32_t s1,
                      float* rgb, uint8_t* found);

/* Parity hooks for RayTracer::rayTrace (source/RayTracer.h:27-53) on caller-supplied rays.
 * flags: 0 = BVH traversal, RT_FLAG_BRUTE_FORCE = O(T) scan. */
int rt_trace_rays(rt_ctx* ctx, const rt_ray* rays, int64_t n, rt_hit* hits, int32_t flags);
/* the boolean use of rayTrace for shadow rays (source/Renderer.cpp:52-55): any hit on (0, +inf) */
int rt_occluded(rt_ctx* ctx, const rt_ray* rays, int64_t n, uint8_t* occluded, int32_t flags);
/* Material::evaluateColorResponse (source/Material.h:25-36): in 9 floats (normal, wi, wo) per item */
int rt_eval_bsdf(rt_ctx* ctx, const rt_material* material, const float* n_wi_wo, int64_t n, float* rgb);

/* RayTracer::hsphereUniformSample (source/RayTracer.h:95-107, maxRayAngle = pi/2) around n normals (3 floats each):
 * item i draws its four words from stream (seed, domain, index0 + i) starting at word 0 (parity hook). */
int rt_eval_hsphere(rt_ctx* ctx, uint64_t seed, uint64_t domain, uint64_t index0, const float* normals, int64_t n,
                    float* directions);

/* Photon map (source/PhotonMap.h:14-50,92-155).  rt_emit_photons traces paths [first_path,
 * first_path+num_paths) of EVERY light (num_paths < 0: all) and returns the stored particles in
 * (light, path) order together with per-light counts, so shards can be concatenated into the list
 * the single-process run produces.  rt_set_photons installs a list: the kd-tree is built on the
 * host with the reference's procedure (source/kdtree.h:60-69) and uploaded. */
int rt_photons_per_light(const rt_ctx* ctx, int32_t* out);
int rt_emit_photons(rt_ctx* ctx, int32_t first_path, int32_t num_paths, rt_photon* out, int64_t capacity,
                    int64_t* per_light_counts, int32_t* depth_histogram20);
int rt_set_photons(rt_ctx* ctx, const rt_photon* photons, int64_t n);
/* The same with the particles staying in DEVICE memory of ctx's GPU (7 floats each, Particle's layout), for the
 * multi-GPU path: every rank emits its share (rt_emit_photons_device: stored particles compacted on the device in
 * (light, path) order, per-light counts to the host), the shards are all-gathered with NCCL where they lie (rank r's
 * shard at gathered_device + r*stride*7), rt_splice_photons_device rearranges them on the device into the order of the
 * single-process list (counts[r*L + l] = particles of light l in rank r's shard), and rt_set_photons_device installs
 * it with ONE device->host copy for the host kd-tree build (none in the device-built exact mode). */
int rt_emit_photons_device(rt_ctx* ctx, int32_t first_path, int32_t num_paths, float* out7_device, int64_t capacity,
                           int64_t* per_light_counts, int32_t* depth_histogram20);
int rt_splice_photons_device(rt_ctx* ctx, const float* gathered_device, int32_t world, int64_t stride,
                             const int64_t* counts, float* out7_device, int64_t capacity, int64_t* total);
int rt_set_photons_device(rt_ctx* ctx, const float* photons7_device, int64_t n);
int rt_build_photon_map(rt_ctx* ctx); /* emit all + set, as Renderer.cpp:209-213 */
int rt_get_photons(rt_ctx* ctx, rt_photon* out, int64_t capacity, int64_t* count);
/* kdtree::knearest (source/kdtree.h:87-107,180-195) for n query points (3 floats each): indices into
 * the kd-ordered node array (see rt_get_kdtree) in the reference's output order, k per query.
 * With RT_FLAG_KNN_EXACT in the context's flags: the exact k nearest by (distance, index) instead. */
int rt_knn(rt_ctx* ctx, const float* queries, int64_t n, int32_t k, int32_t* node_index);
int rt_get_kdtree(rt_ctx* ctx, rt_photon* nodes, int32_t* left, int32_t* right, int32_t* root, int64_t capacity);

/* Which pixels (y*W + x) a shard owns, in the order the wavefront processes them (pure host function,
 * needs no device): call with out = NULL to get the count. */
int rt_shard_pixels(const rt_params* params, int32_t* out, int64_t capacity, int64_t* count);

/* Tooling hook: cudaProfilerStart (on != 0) / cudaProfilerStop, so that `ncu --profile-from-start off` captures
 * exactly the frames a script brackets (scripts/profile_frame.py). */
int rt_profiler_range(int on);

int rt_get_stats(rt_ctx* ctx, rt_stats* out);
int rt_reset_stats(rt_ctx* ctx);
/* flattened BVH for inspection/tests: nodes*16 floats; returns counts through out params */
int rt_get_bvh(rt_ctx* ctx, float* nodes16, int64_t capacity_nodes, int32_t* num_nodes, int32_t* depth);
/* leaf slot -> global triangle index of the context's BVH (num_triangles ints) */
int rt_get_bvh_slots(rt_ctx* ctx, int32_t* slot_triangle, int64_t capacity);
/* The host-side kd-tree builder alone (pure host function, needs no device): photons7_inout holds n particles in
 * emission order on entry and the node array of kdtree::make_tree (source/kdtree.h:60-69, in-order layout) on exit.
 * canonical = 0: libstdc++'s std::nth_element with the reference's comparator -- the reference's own tree;
 * canonical != 0: the tree of the exact k-NN mode (photons ordered by (coordinate, list index)), the one the device
 * builder (csrc/kd_build.cu) produces; orig_index (nullable) receives the list index of every node. */
int rt_build_kdtree_host(float* photons7_inout, int64_t n, int32_t canonical, int32_t* orig_index, int32_t* height);
/* The host-side BVH builder alone (pure host function, needs no device): the split policy of
 * BVH::from_triangles (source/BVH.h:100-161) per mesh plus the top-level join, exactly what rt_create uploads.
 * nodes16: capacity_nodes*16 floats (layout in csrc/host_build.h); slot_triangle: num_triangles ints
 * (leaf slot -> global triangle index).  Call with nodes16 = NULL to get the counts. */
int rt_build_bvh_host(const rt_scene* scene, float pad_fraction, float* nodes16, int64_t capacity_nodes,
                      int32_t* slot_triangle, int32_t* num_nodes, int32_t* depth);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
