// capi.cu -- the C ABI declared in include/rt_b200.h: context, device memory, wavefront scheduling.
//
// No CPU fallback lives here: every entry point needs a CUDA device and fails with RT_ERR_NO_DEVICE /
// RT_ERR_CUDA otherwise.  Nothing under oracle/ is linked or called.
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rt_b200.h"
#include "bvh_build.h"
#include "host_build.h"
#include "kd_build.h"
#include "kernels.h"

using namespace rtb;

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(e_ == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA,                         \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                                \
  } while (0)

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync) with the release
// threshold raised to "never": a context that is destroyed returns its multi-GB wavefront buffers to the
// pool, and the next rt_create on that GPU gets them back without a driver allocation (the e2e path of
// bench.py creates a context per frame).
thread_local cudaStream_t g_alloc_stream = nullptr;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t count) {
    if (count <= n) return cudaSuccess;
    if (p) cudaFreeAsync(p, g_alloc_stream);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMallocAsync((void**)&p, std::max<size_t>(count, 1) * sizeof(T), g_alloc_stream);
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() {
    if (p) cudaFreeAsync(p, g_alloc_stream);
    p = nullptr;
    n = 0;
  }
};

inline float3 h3(const float* a) { return make_float3(a[0], a[1], a[2]); }
inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

struct rt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  rt_params params{};
  int V = 0, T = 0, M = 0, L = 0;
  int num_sms = 148;
  // scene
  DevBuf<float4> d_nodes, d_tris, d_pos, d_nrm;
  DevBuf<int4> d_tri_vidx;
  DevBuf<DMaterial> d_mats;
  DScene scene{};
  Bvh bvh;
  bool bvh_on_device = false;  // built by csrc/bvh_build.cu: bvh.nodes holds only the top-level join on the host
  // photon map
  std::vector<float> kd_nodes7;   // host copy of the kd-ordered nodes (lazily fetched when the tree was built on the device)
  bool kd_host_valid = true;
  int64_t kd_n = 0;
  DevBuf<int> d_kd_orig;           // device-built tree: list index of the photon at every array position
  bool kd_on_device = false;
  DevBuf<float4> d_kd_pos, d_kd_dir;
  int kd_height = 0;
  bool photon_map_built = false;
  uint64_t photon_map_seed = 0;
  int photon_map_requested = -1;
  // wavefront work buffers
  DevBuf<int> d_pix_map;
  int npix = 0;
  int pix_w = -1, pix_h = -1, pix_rank = -1, pix_count = -1, pix_tile = -1;
  DevBuf<float4> d_col0, d_col1, d_col2, d_qo0, d_qo1, d_qd0, d_qd1, d_acc;
  DevBuf<float4> d_hit, d_hit_p, d_sh_d, d_contrib;
  DevBuf<DLight> d_lights_ext;             // lights beyond the kMaxLights kept in kernel-parameter space
  DevBuf<unsigned long long> d_knn_scratch;  // k-NN candidates of every resident thread when k > kKnnSharedMaxK
  int knn_scratch_threads = 0;
  int knn_gather = 0;  // RT_KNN_GATHER=1: the photon queries run in the persistent gather kernel instead of inside k_shade
  int sort_seg0 = 1;   // RT_SORT_SEG0=0: the photon gather of segment 0 keeps the pixel-tile order of the primary hits
  bool tile_force = false;  // RT_SHADE_TILE_FORCE=1: keep the tile size even when the frame has few queries (tests)
  int tile_rounds = 8;  // RT_SHADE_TILE_ROUNDS=r (power of two): the photon k_shade orders each tile of r * 128 slots by fine Morton code
  int sort_bits = -1;  // RT_SORT_BITS=0|5|6|7: Morton cells per axis = 2^bits (0: the 32^3 shared-memory counting sort);
                       // default: 6 with a photon map (k-NN queries gain more from neighbours than the finer sort costs), else 0
  int sort_segs = -1;  // RT_SORT_SEGS=mask: which segments are Morton-binned (default: 1 and 2, plus 0 with a photon map)
  int own_tri = 0;  // RT_OWN_TRI=1: k_shade pre-tests a shadow ray against the triangle it starts on (measured: no gain)
  DevBuf<unsigned char> d_occ;
  DevBuf<int> d_hit_path;
  DevBuf<unsigned int> d_perm, d_sort_hist;
  DevBuf<float4> d_sorted;
  float3 bounds_lo{0, 0, 0}, bounds_hi{0, 0, 0};
  int sort_hits = 1;  // RT_SORT_HITS=0 disables the spatial binning of bounce segments
  DevBuf<int> d_acc_cnt, d_out_cnt;
  DevBuf<float> d_out_rgb;
  DevBuf<unsigned int> d_qcount;
  DevBuf<unsigned long long> d_counters;
  DevBuf<unsigned char> d_scratch;
  size_t auto_paths = 0;  // cached batch-size decision (paths per wavefront batch)
  bool oom_injected = false;  // RT_TEST_OOM_ONCE (tests only)
  bool counters_dirty = false;  // the device ray / query counters moved since they were last read back
  // CUDA-event pairs around the kernel launches of the current render call, by kernel class
  struct TimedSpan {
    int cls;
    cudaEvent_t e0, e1;
  };
  std::vector<TimedSpan> spans;
  size_t spans_used = 0;
  rt_stats stats{};
};

namespace {

int bind(rt_ctx* c) {
  if (!c) return fail(RT_ERR_INVALID, "null context");
  CU(cudaSetDevice(c->device));
  g_alloc_stream = c->stream;
  return RT_OK;
}

// Pixels owned by this shard, tile by tile, 8x4 micro-tiles inside a tile so that the 32 lanes of a
// warp cover a compact block of the image (coherent primary rays).
void build_pix_map(const rt_params& p, std::vector<int>& map) {
  map.clear();
  const int W = p.width, H = p.height;
  const int tile = p.shard_tile > 0 ? p.shard_tile : 16;
  const int count = p.shard_count > 1 ? p.shard_count : 1;
  const int rank = count > 1 ? p.shard_rank : 0;
  const int tx_n = (W + tile - 1) / tile, ty_n = (H + tile - 1) / tile;
  for (int ty = 0; ty < ty_n; ty++)
    for (int tx = 0; tx < tx_n; tx++) {
      if ((ty * tx_n + tx) % count != rank) continue;
      for (int my = 0; my < tile; my += 4)
        for (int mx = 0; mx < tile; mx += 8)
          for (int dy = 0; dy < 4; dy++)
            for (int dx = 0; dx < 8; dx++) {
              int x = tx * tile + mx + dx, y = ty * tile + my + dy;
              if (mx + dx < tile && my + dy < tile && x < W && y < H) map.push_back(y * W + x);
            }
    }
}

// the host side of the map is a pure function of (W, H, rank, count, tile): keep the last few (a context per frame
// in the e2e path would otherwise rebuild 176 k entries every time)
std::vector<int> cached_pix_map(const rt_params& p) {
  struct Entry {
    int w, h, rank, count, tile;
    std::vector<int> map;
  };
  static std::mutex mu;
  static std::vector<Entry> cache;
  const int count = p.shard_count > 1 ? p.shard_count : 1, rank = count > 1 ? p.shard_rank : 0;
  const int tile = p.shard_tile > 0 ? p.shard_tile : 16;
  std::lock_guard<std::mutex> lock(mu);
  for (const Entry& e : cache)
    if (e.w == p.width && e.h == p.height && e.rank == rank && e.count == count && e.tile == tile) return e.map;
  if (cache.size() >= 4) cache.erase(cache.begin());
  cache.push_back(Entry{p.width, p.height, rank, count, tile, {}});
  build_pix_map(p, cache.back().map);
  return cache.back().map;
}

int ensure_pix_map(rt_ctx* c) {
  const rt_params& p = c->params;
  int count = p.shard_count > 1 ? p.shard_count : 1, rank = count > 1 ? p.shard_rank : 0;
  int tile = p.shard_tile > 0 ? p.shard_tile : 16;
  if (c->pix_w == p.width && c->pix_h == p.height && c->pix_rank == rank && c->pix_count == count &&
      c->pix_tile == tile)
    return RT_OK;
  const std::vector<int> map = cached_pix_map(p);
  CU(c->d_pix_map.ensure(map.size()));
  // stream-ordered allocation, copy and consumers share one stream; the source is pageable, so the call returns
  // only once it has been staged (the vector may go away)
  if (!map.empty())
    CU(cudaMemcpyAsync(c->d_pix_map.p, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  c->npix = (int)map.size();
  c->pix_w = p.width;
  c->pix_h = p.height;
  c->pix_rank = rank;
  c->pix_count = count;
  c->pix_tile = tile;
  return RT_OK;
}

// which restatement of kdtree::knearest runs (kernels.h RenderArgs::knn_exact): the ascending array wins for small k
// (cfg3, k = 10: 71 ms per frame against 78 with the heap), the heap for large k (cfg4, k = 50: 3.5 ms against 6.9);
// measured crossover at k ~ 12 (scripts/knn_k_sweep.py)
int knn_flavour_k(const rt_params& p, int k) {
  if (p.flags & RT_FLAG_KNN_EXACT) return 1;
  const char* e = getenv("RT_KNN_HEAP_FROM_K");  // read per call: the tests switch it
  const int heap_from = e ? atoi(e) : 12;
  return (k >= heap_from || k > kMaxK) ? -1 : 0;  // the ascending array's tie fallback holds at most kMaxK candidates
}
int knn_flavour(const rt_params& p) { return knn_flavour_k(p, p.k); }

// stack entries a traversal can need: the whole tree from node 0 (k_emit), or the root list plus one mesh's tree
int trace_stack_depth(const rt_ctx* c) {
  return std::max(std::max(c->bvh.depth, c->scene.num_roots > 0 ? c->bvh.mesh_depth + c->scene.num_roots : 0), 1);
}

// bytes of wavefront state per path slot (ensure_work below) with nl shadow slots per hit: colours 3 x 16, ray queues
// 2 x 32, hit record 16, hit point 16, per shadow slot direction 16 + contribution 16 + occlusion 1, path id 4,
// sort key 4, sorted payload 32
size_t bytes_per_path(int nl) { return 16 * 3 + 32 * 2 + 16 + 16 + (size_t)std::max(nl, 1) * (16 + 16 + 1) + 4 + 4 + 32; }

void release_work(rt_ctx* c) {
  c->d_col0.release();
  c->d_col1.release();
  c->d_col2.release();
  c->d_qo0.release();
  c->d_qo1.release();
  c->d_qd0.release();
  c->d_qd1.release();
  c->d_hit.release();
  c->d_hit_path.release();
  c->d_hit_p.release();
  c->d_sh_d.release();
  c->d_contrib.release();
  c->d_occ.release();
  c->d_perm.release();
  c->d_sorted.release();
}

// shadow slots per hit: one per light for direct lighting, one (the gathered contribution) with a photon map
int shadow_lights(const rt_ctx* c, bool use_photons) { return use_photons ? 1 : c->L; }

int ensure_work(rt_ctx* c, size_t paths, bool path_mode, int nl) {
  const size_t shadow = std::max<size_t>(shadow_slots_for(paths, (unsigned)std::max(nl, 0)), 1);
  if (shadow >= ((size_t)1 << 32)) return fail(RT_ERR_INVALID, "batch too large for the 32-bit shadow-ray queue");
  CU(c->d_col0.ensure(paths));
  CU(c->d_qo0.ensure(paths));
  CU(c->d_qd0.ensure(paths));
  CU(c->d_hit.ensure(paths));
  CU(c->d_hit_path.ensure(paths));
  CU(c->d_hit_p.ensure(paths));
  CU(c->d_sh_d.ensure(shadow));
  CU(c->d_contrib.ensure(shadow));
  CU(c->d_occ.ensure(shadow));
  if (path_mode) {
    CU(c->d_col1.ensure(paths));
    CU(c->d_col2.ensure(paths));
    CU(c->d_qo1.ensure(paths));
    CU(c->d_qd1.ensure(paths));
  }
  if (path_mode || c->params.num_photons > 0) {
    CU(c->d_perm.ensure(paths));
    CU(c->d_sorted.ensure(2 * paths));
    const bool fine = c->sort_bits > 0 || (c->sort_bits < 0 && c->params.num_photons > 0);
    CU(c->d_sort_hist.ensure(fine ? (size_t)kSortFineMax + 2 + kSortFineTiles : (size_t)kSortBuckets + 2));
  }
  CU(c->d_qcount.ensure(kQNum));
  CU(c->d_counters.ensure(kCntNum));
  return RT_OK;
}

int validate_params(const rt_params* p) {
  if (!p) return fail(RT_ERR_INVALID, "null params");
  if (p->width < 1 || p->height < 1) return fail(RT_ERR_INVALID, "width/height must be >= 1");
  if (p->num_rays < 0) return fail(RT_ERR_INVALID, "num_rays must be >= 0");
  if (p->num_photons < 0) return fail(RT_ERR_INVALID, "num_photons must be >= 0");
  if (p->num_photons > 0 && (p->k < 1 || p->k > RT_MAX_K))
    return fail(RT_ERR_INVALID, "k must be in [1, " + std::to_string(RT_MAX_K) + "] when a photon map is used");
  if (p->num_photons >= (1 << 28)) return fail(RT_ERR_INVALID, "too many photons");
  if (p->shard_count > 1 && (p->shard_rank < 0 || p->shard_rank >= p->shard_count))
    return fail(RT_ERR_INVALID, "shard_rank out of range");
  if ((int64_t)p->width * p->height > (int64_t)1 << 30) return fail(RT_ERR_INVALID, "image too large");
  if (p->sample_first < 0 || (p->sample_count > 0 && p->sample_first + p->sample_count > p->num_rays))
    return fail(RT_ERR_INVALID, "sample range outside [0, num_rays)");
  return RT_OK;
}

// k > kKnnSharedMaxK: the candidates of every resident thread of the k-NN shade kernel live in global memory
int ensure_knn_scratch(rt_ctx* c, int k, int threads) {
  if (k <= kKnnSharedMaxK) return RT_OK;
  CU(c->d_knn_scratch.ensure((size_t)k * (size_t)threads));
  c->knn_scratch_threads = threads;
  return RT_OK;
}

void fill_args(rt_ctx* c, RenderArgs& a, const int* pix_map, int npix, bool use_photons) {
  const rt_params& p = c->params;
  a.scene = c->scene;
  a.width = p.width;
  a.height = p.height;
  a.num_rays = p.num_rays;
  a.jitter_d = (int)sqrtf((float)p.num_rays);  // RayTracer.h:111
  if (a.jitter_d < 1) a.jitter_d = 1;
  a.mode = p.mode == 1 ? 1 : 0;                // CommandLine.h:84-87: anything else is ray tracing
  a.photon = use_photons ? 1 : 0;
  a.k = p.k;
  a.knn_exact = knn_flavour(p);
  a.kd_frames = c->kd_height + 1;
  a.num_sms = c->num_sms;
  a.num_photons = p.num_photons;
  a.brute = (p.flags & RT_FLAG_BRUTE_FORCE) ? 1 : 0;
  a.seed_mixed = mix64(p.seed + kGolden);
  a.pix_map = pix_map;
  a.npix = npix;
  a.stack_depth = trace_stack_depth(c);
  a.col0 = c->d_col0.p;
  a.col1 = c->d_col1.p;
  a.col2 = c->d_col2.p;
  a.ray_o[0] = c->d_qo0.p;
  a.ray_o[1] = c->d_qo1.p;
  a.ray_d[0] = c->d_qd0.p;
  a.ray_d[1] = c->d_qd1.p;
  a.hit = c->d_hit.p;
  a.nl = shadow_lights(c, use_photons);
  a.own_tri = c->own_tri;
  a.sort_mask = c->sort_segs >= 0 ? c->sort_segs : (6 | ((use_photons && c->sort_seg0) ? 1 : 0));
  a.hit_p = c->d_hit_p.p;
  a.sh_d = c->d_sh_d.p;
  // the persistent gather handles the two reference-exact flavours; its result array reuses the (unused in photon
  // mode) shadow-ray direction buffer: one float4 per ray slot
  a.knn_out = (use_photons && c->knn_gather && a.knn_exact <= 0) ? c->d_sh_d.p : nullptr;
  a.knn_scratch = c->d_knn_scratch.p;
  a.knn_scratch_stride = c->knn_scratch_threads;
  a.contrib = c->d_contrib.p;
  a.occ = c->d_occ.p;
  a.hit_path = c->d_hit_path.p;
  a.perm = (c->sort_hits && (p.mode == 1 || use_photons)) ? c->d_perm.p : nullptr;
  a.sort_hist = c->d_sort_hist.p;
  a.sorted = c->d_sorted.p;
  a.sort_lo = c->bounds_lo;
  a.sort_bits = c->sort_bits >= 0 ? c->sort_bits : (use_photons ? 6 : 0);
  // the tile sort borrows the shared memory of the queries it precedes (8 bytes per slot of a tile; tiles are powers
  // of two for the bitonic network) and parks its order in the perm array
  a.tile_rounds = 0;
  if (use_photons && c->tile_rounds > 1 && a.knn_out == nullptr && a.perm != nullptr) {
    int rounds = 1;
    while (2 * rounds <= c->tile_rounds && knn_smem_bytes(p.k, c->kd_height + 1) >= (size_t)2 * rounds * kBlock * 8) rounds *= 2;
    a.tile_rounds = rounds > 1 ? (c->tile_force ? -rounds : rounds) : 0;
  }
  const float cells = a.sort_bits > 0 ? (float)(1 << a.sort_bits) : (float)kSortGrid;
  a.sort_inv_cell = make_float3(cells / std::max(c->bounds_hi.x - c->bounds_lo.x, 1e-20f),
                                cells / std::max(c->bounds_hi.y - c->bounds_lo.y, 1e-20f),
                                cells / std::max(c->bounds_hi.z - c->bounds_lo.z, 1e-20f));
  a.sort_key_scale = 1024.f / cells;
  a.q_count = c->d_qcount.p;
  a.counters = c->d_counters.p;
}

// CUDA-event pair around the launches of one kernel class (on the launching stream)
int span_begin(rt_ctx* c, int cls) {
  if (c->spans.size() <= c->spans_used) {
    rt_ctx::TimedSpan sp{cls, nullptr, nullptr};
    CU(cudaEventCreate(&sp.e0));
    CU(cudaEventCreate(&sp.e1));
    c->spans.push_back(sp);
  }
  c->spans[c->spans_used].cls = cls;
  CU(cudaEventRecord(c->spans[c->spans_used].e0, c->stream));
  return RT_OK;
}
int span_end(rt_ctx* c, int launches) {
  CU(cudaEventRecord(c->spans[c->spans_used].e1, c->stream));
  c->stats.kernel_count[c->spans[c->spans_used].cls] += (uint64_t)launches;
  c->stats.kernel_launches += (uint64_t)launches;
  c->spans_used++;
  return RT_OK;
}
// after the stream has been synchronised: add the spans' times to the per-class totals (cumulative since
// rt_reset_stats); trace_ms is the trace time of THIS call
int collect_spans(rt_ctx* c) {
  double trace = 0.0;
  for (size_t i = 0; i < c->spans_used; i++) {
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, c->spans[i].e0, c->spans[i].e1));
    c->stats.kernel_ms[c->spans[i].cls] += t;
    if (c->spans[i].cls == kKTraceNearest || c->spans[i].cls == kKTraceAny) trace += t;
  }
  c->spans_used = 0;
  c->stats.trace_ms = trace;
  return RT_OK;
}
void reset_call_timing(rt_ctx* c) { c->spans_used = 0; }

// one wavefront batch: samples [s0, s0+nsamp) of the pixels in pix_map
int run_batch(rt_ctx* c, RenderArgs& a, int s0, int nsamp) {
  a.s0 = s0;
  a.nsamp = nsamp;
  const long long paths = (long long)a.npix * nsamp;
  const int persistent = c->num_sms * trace_ctas_per_sm(a.stack_depth);
  int grid = (int)std::min<long long>((paths + kBlock - 1) / kBlock, (long long)persistent);
  if (grid < 1) grid = 1;
  // the any-hit kernel keeps no entry-distance column: its own (higher) occupancy sizes its persistent grid
  int grid_any = (int)std::min<long long>((paths * std::max(a.nl, 1) + kBlock - 1) / kBlock,
                                          (long long)c->num_sms * trace_any_ctas_per_sm(a.stack_depth));
  if (grid_any < 1) grid_any = 1;
  const int grid_shade = (int)std::min<long long>((paths + kBlock - 1) / kBlock, (long long)c->num_sms * 8);
  CU(cudaMemsetAsync(c->d_qcount.p, 0, kQNum * sizeof(unsigned), c->stream));
  int rc;
#define SPAN(cls, launches, call)          \
  do {                                     \
    if ((rc = span_begin(c, cls))) return rc; \
    call;                                  \
    if ((rc = span_end(c, launches))) return rc; \
  } while (0)
  SPAN(kKRaygen, 1, launch_raygen(a, c->stream));
  const int nseg = a.mode == 1 ? 3 : 1;
  for (int seg = 0; seg < nseg; seg++) {
    SPAN(kKTraceNearest, 1, launch_trace_nearest(a, seg, grid, c->stream));
    if (((a.sort_mask >> seg) & 1) && a.perm)  // spatial order: shadow-ray / k-NN traversal coherence
      SPAN(kKSort, 3, launch_sort_hits(a, seg, c->stream));
    if (a.photon && a.knn_out) SPAN(kKGather, 1, launch_knn_gather(a, seg, c->stream));
    SPAN(kKShade, 1, launch_shade(a, seg, std::max(grid_shade, 1), c->stream));
    if (!a.photon && a.nl > 0) SPAN(kKTraceAny, 1, launch_trace_any(a, seg, grid_any, c->stream));
    if (!a.photon) SPAN(kKCombine, 1, launch_combine(a, seg, std::max(grid_shade, 1), c->stream));
  }
#undef SPAN
  CU(cudaGetLastError());
  return RT_OK;
}

// The device counters are read back only when somebody asks for the statistics (rt_get_stats): a render call does
// not pay a device->host round trip for them.
int pull_counters(rt_ctx* c) {
  c->counters_dirty = true;
  return RT_OK;
}
int read_counters(rt_ctx* c) {
  if (!c->counters_dirty) return RT_OK;
  c->counters_dirty = false;
  unsigned long long h[kCntNum];
  CU(cudaMemcpyAsync(h, c->d_counters.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  rt_stats& s = c->stats;
  s.shadow_rays = h[kCntShadow];
  s.photon_rays = h[kCntPhotonRays];
  s.knn_queries = h[kCntKnn];
  s.kd_visits = h[kCntKdVisits];
  uint64_t nearest = h[kCntNearest];
  s.bounce_rays = nearest >= s.samples ? nearest - s.samples : 0;
  s.primary_rays = nearest - s.bounce_rays;
  s.rays = nearest + s.shadow_rays + s.photon_rays;
  return RT_OK;
}

int photons_per_light(const rt_ctx* c, float* light_pdf_out) {
  // PhotonMap.h:19-20: float lightPdf = 1.f / size; int photonsPerLS = (int)(numOfPhotons * lightPdf)
  if (c->L < 1) return 0;
  float light_pdf = 1.f / (float)c->L;
  if (light_pdf_out) *light_pdf_out = light_pdf;
  return (int)((float)c->params.num_photons * light_pdf);
}

// fn(begin, end) over [0, n) on up to 8 host threads when n is large (scene packing for million-triangle scenes)
template <typename F>
void parallel_ranges(int64_t n, F fn) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = n >= (1 << 18) ? (int)std::min<unsigned>(8, hw ? hw : 1) : 1;
  if (nt <= 1) {
    fn((int64_t)0, n);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++) th.emplace_back([=]() { fn(n * t / nt, n * (t + 1) / nt); });
  for (auto& x : th) x.join();
}

// largest |coordinate| of any vertex, light or the camera: scales the BVH padding (host_build.h)
float scene_extent(const rt_scene* s) {
  float extent = 0.f;
  for (int64_t i = 0; i < 3 * (int64_t)s->num_vertices; i++) {
    const float a = std::fabs(s->positions[i]);
    if (!(a <= FLT_MAX)) return INFINITY;  // NaN or inf: std::max would silently skip a NaN
    extent = std::max(extent, a);
  }
  for (int l = 0; l < s->num_lights; l++)
    for (int a = 0; a < 3; a++) extent = std::max(extent, std::fabs(s->lights[l].position[a]) + s->lights[l].side);
  for (int a = 0; a < 3; a++) extent = std::max(extent, std::fabs(s->camera.position[a]));
  return extent;
}

bool use_photon_map(const rt_ctx* c) { return c->params.num_photons > 0 && c->scene.kd_count > 0; }

int prepare_photons(rt_ctx* c) {
  const rt_params& p = c->params;
  if (p.num_photons <= 0) return RT_OK;
  if (!c->photon_map_built || c->photon_map_seed != p.seed || c->photon_map_requested != p.num_photons) {
    int rc = rt_build_photon_map(c);
    if (rc != RT_OK) return rc;
  }
  // Renderer.cpp:239,245: an empty tree silently selects the direct-lighting overloads
  if (c->scene.kd_count > 0 && p.k > c->scene.kd_count)
    return fail(RT_ERR_K_TOO_LARGE, "k is greater than the number of nodes");  // kdtree.h:182-183
  return RT_OK;
}

// wavefront buffers (and the k-NN scratch) for `paths` path slots, with the batch-size fallback of the auto mode
int prepare_batch_buffers(rt_ctx* c, bool use_photons) {
  if (!use_photons || c->params.k <= kKnnSharedMaxK) return RT_OK;
  const int flavour = knn_flavour(c->params);
  int per_sm = shade_photon_ctas_per_sm(c->params.mode == 1 ? 1 : 0, c->params.k, c->kd_height + 1);
  if (flavour <= 0) per_sm = std::max(per_sm, knn_gather_ctas_per_sm(c->params.k, c->kd_height + 1, flavour < 0 ? 1 : 0));
  return ensure_knn_scratch(c, c->params.k, c->num_sms * per_sm * kBlock);
}

// the whole render: batches of samples -> ordered accumulation -> scatter into full-frame buffers
// composite_dev != null: instead of the raw sums/counters, composite over the background it holds (rt_render)
// packed_dev != null: the sums and the counter as one float4 per pixel (counter exact as a float up to 2^24 samples)
struct Progress {
  int every = 0;
  rt_progress_fn fn = nullptr;
  void* user = nullptr;
};
// background_host != null: the host frame composite_dev is to be filled from.  It is uploaded AFTER the sample batches have
// been enqueued (the staging of a pageable source then runs on the host while the GPU traces), or up front when a
// preview needs it between batches.
int render_to_device(rt_ctx* c, float* out_rgb_dev, int* out_cnt_dev, float* composite_dev = nullptr,
                     const Progress* progress = nullptr, float4* packed_dev = nullptr,
                     const float* background_host = nullptr) {
  static const bool timing = getenv("RT_TIMING") != nullptr;
  double t_prev = now_ms();
  auto tick = [&](const char* what) {
    if (!timing) return;
    const double t = now_ms();
    fprintf(stderr, "[render] %s %.3f ms; ", what, t - t_prev);
    t_prev = t;
  };
  int rc = bind(c);
  if (rc) return rc;
  reset_call_timing(c);
  if ((rc = prepare_photons(c))) return rc;
  if ((rc = ensure_pix_map(c))) return rc;
  tick("bind+pixmap");
  const rt_params& p = c->params;
  const size_t npx = (size_t)p.width * p.height;
  const bool path_mode = p.mode == 1;
  const bool photons = use_photon_map(c);
  const int nl = shadow_lights(c, photons);
  if ((rc = prepare_batch_buffers(c, photons))) return rc;
  // Batch size: <= 32 M paths (bytes_per_path() of wavefront state each, 8.9 GB with three lights) by default.  Free HBM is
  // only asked for when that allocation fails (cudaMemGetInfo costs 1.2-1.6 ms, a twentieth of a cfg2 frame): then
  // the batch shrinks to half of what is free and the allocation is retried once.
  int spb = p.samples_per_batch;
  const bool auto_batch = spb <= 0;
  const size_t max_paths = std::min<size_t>((size_t)32 << 20, (((size_t)1 << 32) - 64) / (size_t)std::max(nl, 1));
  if (auto_batch) {
    if (c->auto_paths == 0) c->auto_paths = max_paths;
    c->auto_paths = std::min(c->auto_paths, max_paths);
    spb = (int)std::max<size_t>(1, std::min<size_t>(c->auto_paths / std::max(c->npix, 1), 1 << 20));
  }
  const int samp_first = p.sample_first;
  const int samp_end = p.sample_count > 0 ? p.sample_first + p.sample_count : p.num_rays;
  const bool preview = progress && progress->fn && progress->every > 0 && composite_dev;
  if (preview) spb = std::min(spb, progress->every);  // a snapshot needs a batch boundary
  spb = std::max(1, std::min(spb, std::max(samp_end - samp_first, 1)));
  std::vector<float> snapshot;
  if (preview) {
    snapshot.resize(3 * npx);
    CU(c->d_scratch.ensure(sizeof(float) * 3 * npx));
  }
  bool background_up = !(background_host && composite_dev);
  if (!background_up && preview) {
    CU(cudaMemcpyAsync(composite_dev, background_host, sizeof(float) * 3 * npx, cudaMemcpyHostToDevice, c->stream));
    background_up = true;
  }
  tick("batch sizing");
  rc = ensure_work(c, (size_t)c->npix * spb, path_mode, nl);
  if (auto_batch && rc == RT_OK && getenv("RT_TEST_OOM_ONCE") && !c->oom_injected) {  // test hook for the retry path
    c->oom_injected = true;
    rc = RT_ERR_OOM;
  }
  if (rc == RT_ERR_OOM && auto_batch) {
    cudaGetLastError();
    release_work(c);
    CU(cudaStreamSynchronize(c->stream));
    {  // hand the pool's cached blocks back so that the query below sees them as free
      cudaMemPool_t pool;
      CU(cudaDeviceGetDefaultMemPool(&pool, c->device));
      CU(cudaMemPoolTrimTo(pool, 0));
    }
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (getenv("RT_TEST_OOM_ONCE")) free_b = std::min<size_t>(free_b, (size_t)16 << 20);  // force several batches
    c->auto_paths = std::max<size_t>(1, std::min<size_t>(free_b / 2 / bytes_per_path(nl), max_paths));
    spb = (int)std::max<size_t>(1, std::min<size_t>(c->auto_paths / std::max(c->npix, 1), 1 << 20));
    spb = std::max(1, std::min(spb, std::max(samp_end - samp_first, 1)));
    if (preview) spb = std::max(1, std::min(spb, progress->every));
    rc = ensure_work(c, (size_t)c->npix * spb, path_mode, nl);
  }
  if (rc) return rc;
  tick("ensure_work");
  CU(c->d_acc.ensure(c->npix));
  CU(c->d_acc_cnt.ensure(c->npix));
  CU(cudaMemsetAsync(c->d_acc.p, 0, sizeof(float4) * (size_t)c->npix, c->stream));
  CU(cudaMemsetAsync(c->d_acc_cnt.p, 0, sizeof(int) * (size_t)c->npix, c->stream));
  RenderArgs a;
  fill_args(c, a, c->d_pix_map.p, c->npix, photons);
  CU(cudaEventRecord(c->ev0, c->stream));
  if (c->npix > 0) {
    for (int s0 = samp_first; s0 < samp_end; s0 += spb) {
      int ns = std::min(spb, samp_end - s0);
      if ((rc = run_batch(c, a, s0, ns))) return rc;
      if ((rc = span_begin(c, kKResolve))) return rc;
      launch_resolve(c->d_col0.p, c->d_col1.p, c->d_col2.p, path_mode ? 1 : 0, c->npix, ns, c->d_acc.p, c->d_acc_cnt.p,
                     c->stream);
      if ((rc = span_end(c, 1))) return rc;
      const int done = s0 + ns - samp_first;
      if (preview && s0 + ns < samp_end && done / progress->every != (done - ns) / progress->every) {
        // Renderer.cpp:262-269 after pass i = done-1: the composite of the first `done` samples
        float* snap_dev = reinterpret_cast<float*>(c->d_scratch.p);
        CU(cudaMemcpyAsync(snap_dev, composite_dev, sizeof(float) * 3 * npx, cudaMemcpyDeviceToDevice, c->stream));
        launch_composite(c->d_acc.p, c->d_acc_cnt.p, c->d_pix_map.p, c->npix, done, composite_dev, snap_dev, c->stream);
        c->stats.kernel_launches++;
        CU(cudaMemcpyAsync(snapshot.data(), snap_dev, sizeof(float) * 3 * npx, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        progress->fn(progress->user, done, samp_end - samp_first, snapshot.data());
      }
    }
  }
  if (!background_up)
    CU(cudaMemcpyAsync(composite_dev, background_host, sizeof(float) * 3 * npx, cudaMemcpyHostToDevice, c->stream));
  if ((rc = span_begin(c, kKOther))) return rc;
  int tail_launches = 0;
  if (composite_dev) {
    if (c->npix > 0) {
      launch_composite(c->d_acc.p, c->d_acc_cnt.p, c->d_pix_map.p, c->npix, p.num_rays, composite_dev, composite_dev,
                       c->stream);
      tail_launches++;
    }
  } else if (packed_dev) {
    CU(cudaMemsetAsync(packed_dev, 0, sizeof(float4) * npx, c->stream));
    if (c->npix > 0) {
      launch_scatter_packed(c->d_acc.p, c->d_acc_cnt.p, c->d_pix_map.p, c->npix, packed_dev, c->stream);
      tail_launches++;
    }
  } else {
    CU(cudaMemsetAsync(out_rgb_dev, 0, sizeof(float) * 3 * npx, c->stream));
    CU(cudaMemsetAsync(out_cnt_dev, 0, sizeof(int) * npx, c->stream));
    if (c->npix > 0) {
      launch_scatter(c->d_acc.p, c->d_acc_cnt.p, c->d_pix_map.p, c->npix, out_rgb_dev, out_cnt_dev, c->stream);
      tail_launches++;
    }
  }
  if ((rc = span_end(c, tail_launches))) return rc;
  CU(cudaEventRecord(c->ev1, c->stream));
  tick("launches");
  CU(cudaStreamSynchronize(c->stream));
  tick("sync");
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->stats.device_ms = ms;
  c->stats.device_ms_total += ms;
  if ((rc = collect_spans(c))) return rc;
  c->stats.samples += (uint64_t)c->npix * (uint64_t)std::max(samp_end - samp_first, 0);
  rc = pull_counters(c);
  tick("events+counters");
  if (timing) fprintf(stderr, "\n");
  return rc;
}

}  // namespace

extern "C" {

const char* rt_last_error(void) { return g_err.c_str(); }

int rt_abi_sizes(int32_t* out, int32_t capacity) {
  const int32_t sizes[9] = {(int32_t)sizeof(rt_material), (int32_t)sizeof(rt_light),  (int32_t)sizeof(rt_camera),
                            (int32_t)sizeof(rt_scene),    (int32_t)sizeof(rt_params), (int32_t)sizeof(rt_ray),
                            (int32_t)sizeof(rt_hit),      (int32_t)sizeof(rt_photon), (int32_t)sizeof(rt_stats)};
  if (!out || capacity < 9) return fail(RT_ERR_INVALID, "rt_abi_sizes needs room for 9 values");
  for (int i = 0; i < 9; i++) out[i] = sizes[i];
  return RT_OK;
}

int rt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int rt_destroy(rt_ctx* c) {
  if (!c) return RT_OK;
  cudaSetDevice(c->device);
  g_alloc_stream = c->stream;
  c->d_nodes.release();
  c->d_tris.release();
  c->d_pos.release();
  c->d_nrm.release();
  c->d_tri_vidx.release();
  c->d_mats.release();
  c->d_kd_pos.release();
  c->d_kd_dir.release();
  c->d_pix_map.release();
  c->d_col0.release();
  c->d_col1.release();
  c->d_col2.release();
  c->d_qo0.release();
  c->d_qo1.release();
  c->d_hit.release();
  c->d_hit_p.release();
  c->d_sh_d.release();
  c->d_contrib.release();
  c->d_occ.release();
  c->d_hit_path.release();
  c->d_lights_ext.release();
  c->d_knn_scratch.release();
  c->d_kd_orig.release();
  c->d_perm.release();
  c->d_sorted.release();
  c->d_sort_hist.release();
  c->d_qd0.release();
  c->d_qd1.release();
  c->d_acc.release();
  c->d_acc_cnt.release();
  c->d_out_cnt.release();
  c->d_out_rgb.release();
  c->d_qcount.release();
  c->d_counters.release();
  c->d_scratch.release();
  for (auto& sp : c->spans) {
    cudaEventDestroy(sp.e0);
    cudaEventDestroy(sp.e1);
  }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) {
    cudaStreamSynchronize(c->stream);  // the stream-ordered frees above complete before the stream goes away
    cudaStreamDestroy(c->stream);
  }
  g_alloc_stream = nullptr;
  delete c;
  return RT_OK;
}

int rt_set_params(rt_ctx* c, const rt_params* p) {
  if (!c) return fail(RT_ERR_INVALID, "null context");
  int rc = validate_params(p);
  if (rc) return rc;
  c->params = *p;
  return RT_OK;
}

int rt_create(const rt_scene* s, const rt_params* p, int device, rt_ctx** out) {
  if (!s || !out) return fail(RT_ERR_INVALID, "null scene/out");
  *out = nullptr;
  int rc = validate_params(p);
  if (rc) return rc;
  if (s->num_vertices < 0 || s->num_triangles < 0 || s->num_meshes < 0 || s->num_lights < 0)
    return fail(RT_ERR_INVALID, "negative scene counts");
  if (s->num_lights > 4096) return fail(RT_ERR_INVALID, "more than 4096 light sources");
  if (s->num_triangles > 0 && (!s->positions || !s->normals || !s->triangles || !s->mesh_first_triangle ||
                               !s->mesh_first_vertex || !s->materials))
    return fail(RT_ERR_INVALID, "null scene arrays");
  if (s->num_lights > 0 && !s->lights) return fail(RT_ERR_INVALID, "null lights");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    return fail(RT_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return fail(RT_ERR_INVALID, "device ordinal out of range");
  {
    std::atomic<bool> bad{false};
    parallel_ranges(3 * (int64_t)s->num_triangles, [&](int64_t b, int64_t e) {
      for (int64_t t = b; t < e; t++)
        if (s->triangles[t] < 0 || s->triangles[t] >= s->num_vertices) bad = true;
    });
    if (bad) return fail(RT_ERR_INVALID, "triangle vertex index out of range");
  }

  const double t_create0 = now_ms();
  rt_ctx* c = new rt_ctx();
  c->device = device;
  c->params = *p;
  c->V = s->num_vertices;
  c->T = s->num_triangles;
  c->M = s->num_meshes;
  c->L = s->num_lights;
#define CUC(call)                                                                   \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_);          \
      rt_destroy(c);                                                                \
      return fail(e_ == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA, m_);  \
    }                                                                               \
  } while (0)
  CUC(cudaSetDevice(device));
  CUC(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  g_alloc_stream = c->stream;
  {
    cudaMemPool_t pool;
    CUC(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t never = UINT64_MAX;
    CUC(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &never));
  }
  CUC(cudaEventCreate(&c->ev0));
  CUC(cudaEventCreate(&c->ev1));

  // ---- BVH (host, reference split policy) ----
  const float extent = scene_extent(s);
  if (c->V > 0) {
    float lo[3] = {s->positions[0], s->positions[1], s->positions[2]}, hi[3] = {lo[0], lo[1], lo[2]};
    for (int v = 0; v < c->V; v++)
      for (int a = 0; a < 3; a++) {
        lo[a] = std::min(lo[a], s->positions[3 * v + a]);
        hi[a] = std::max(hi[a], s->positions[3 * v + a]);
      }
    c->bounds_lo = make_float3(lo[0], lo[1], lo[2]);
    c->bounds_hi = make_float3(hi[0], hi[1], hi[2]);
  }
  if (const char* e = getenv("RT_SORT_HITS")) c->sort_hits = atoi(e);
  if (const char* e = getenv("RT_OWN_TRI")) c->own_tri = atoi(e) != 0;
  if (const char* e = getenv("RT_KNN_GATHER")) c->knn_gather = atoi(e) != 0;
  if (const char* e = getenv("RT_SORT_SEG0")) c->sort_seg0 = atoi(e) != 0;
  if (const char* e = getenv("RT_SORT_SEGS")) c->sort_segs = atoi(e) & 7;
  if (const char* e = getenv("RT_SHADE_TILE_FORCE")) c->tile_force = atoi(e) != 0;
  if (const char* e = getenv("RT_SHADE_TILE_ROUNDS")) c->tile_rounds = std::min(std::max(atoi(e), 0), 64);
  if (const char* e = getenv("RT_SORT_BITS")) c->sort_bits = std::min(std::max(atoi(e), 0), 7);
  if (!(extent < 1e8f)) {  // keeps lo * safe_inv(d) finite in the slab test (rt_device.cuh)
    rt_destroy(c);
    return fail(RT_ERR_INVALID, "scene coordinates must be finite and smaller than 1e8");
  }
  float pad_fraction = p->bvh_pad > 0.f ? p->bvh_pad : 1.0f / 16384.0f;

  // ---- shading data up first: the device BVH builder reads positions and triangle indices from HBM ----
  std::vector<float4> h_pos(std::max(c->V, 1)), h_nrm(std::max(c->V, 1));
  std::vector<int4> h_vidx(std::max(c->T, 1));
  parallel_ranges(c->V, [&](int64_t b, int64_t e) {
    for (int64_t v = b; v < e; v++) {
      h_pos[v] = make_float4(s->positions[3 * v], s->positions[3 * v + 1], s->positions[3 * v + 2], 0.f);
      h_nrm[v] = make_float4(s->normals[3 * v], s->normals[3 * v + 1], s->normals[3 * v + 2], 0.f);
    }
  });
  parallel_ranges(c->T, [&](int64_t b, int64_t e) {
    int m = 0;
    for (int64_t t = b; t < e; t++) {
      while (m + 1 < c->M && t >= s->mesh_first_triangle[m + 1]) m++;
      h_vidx[t] = make_int4(s->triangles[3 * t], s->triangles[3 * t + 1], s->triangles[3 * t + 2], m);
    }
  });
  std::vector<DMaterial> h_mats(std::max(c->M, 1));
  for (int m = 0; m < c->M; m++) {
    const rt_material& a = s->materials[m];
    h_mats[m] = make_material(a.kd, a.alpha, h3(a.albedo), h3(a.f0));
  }
  CUC(c->d_pos.ensure(h_pos.size()));
  CUC(c->d_nrm.ensure(h_nrm.size()));
  CUC(c->d_tri_vidx.ensure(h_vidx.size()));
  CUC(c->d_mats.ensure(h_mats.size()));
  // Allocation (stream-ordered pool), upload and the kernels that read these buffers all run on c->stream, so they
  // are ordered without a device-wide sync; pageable sources are staged before cudaMemcpyAsync returns.
  CUC(cudaMemcpyAsync(c->d_pos.p, h_pos.data(), h_pos.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  CUC(cudaMemcpyAsync(c->d_tri_vidx.p, h_vidx.data(), h_vidx.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
  CUC(cudaMemcpyAsync(c->d_nrm.p, h_nrm.data(), h_nrm.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  CUC(cudaMemcpyAsync(c->d_mats.p, h_mats.data(), h_mats.size() * sizeof(DMaterial), cudaMemcpyHostToDevice, c->stream));

  // ---- BVH (reference split policy): on the device for large scenes, on the host otherwise ----
  const double t_bvh0 = now_ms();
  {
    const char* e = getenv("RT_BVH_BUILD");  // "gpu" | "host"; default: gpu from 8192 triangles
    const bool want_gpu = e ? std::string(e) == "gpu" : c->T >= 8192;
    if (want_gpu && c->T >= 2) {
      int nonempty = 0;
      for (int m = 0; m < c->M; m++) nonempty += s->mesh_first_triangle[m + 1] > s->mesh_first_triangle[m] ? 1 : 0;
      const size_t num_nodes = (size_t)(c->T - nonempty) + (size_t)std::max(nonempty - 1, 0);
      CUC(c->d_nodes.ensure(4 * std::max<size_t>(num_nodes, 1)));
      CUC(c->d_tris.ensure(3 * (size_t)c->T));
      std::string err;
      long long launches = 0;
      if (build_bvh_device(c->d_pos.p, c->d_tri_vidx.p, c->T, c->M, s->mesh_first_triangle, extent * pad_fraction,
                           c->d_nodes.p, c->d_tris.p, c->stream, c->bvh, &launches, err)) {
        c->bvh_on_device = true;
        c->stats.kernel_launches += (uint64_t)launches;
      } else if (e) {  // explicitly requested: report instead of silently building on the host
        rt_destroy(c);
        return fail(RT_ERR_CUDA, "device BVH build failed: " + err);
      }
    }
  }
  if (!c->bvh_on_device) {
    build_bvh(c->V, s->positions, c->T, s->triangles, c->M, s->mesh_first_triangle, extent, pad_fraction, c->bvh);
    std::vector<float4> h_tris(3 * (size_t)std::max(c->T, 1));
    for (int slot = 0; slot < c->T; slot++) {
      int gid = c->bvh.slot_tri[slot];
      const float* p0 = s->positions + 3 * (size_t)s->triangles[3 * gid];
      const float* p1 = s->positions + 3 * (size_t)s->triangles[3 * gid + 1];
      const float* p2 = s->positions + 3 * (size_t)s->triangles[3 * gid + 2];
      // Ray.cpp:11: edge1 = p1 - p0, edge2 = p2 - p0 (binary32 subtractions, done once here)
      volatile float e1x = p1[0] - p0[0], e1y = p1[1] - p0[1], e1z = p1[2] - p0[2];
      volatile float e2x = p2[0] - p0[0], e2y = p2[1] - p0[1], e2z = p2[2] - p0[2];
      int gid_bits = gid;
      float gid_f;
      std::memcpy(&gid_f, &gid_bits, 4);
      h_tris[3 * (size_t)slot] = make_float4(p0[0], p0[1], p0[2], gid_f);
      h_tris[3 * (size_t)slot + 1] = make_float4(e1x, e1y, e1z, 0.f);
      h_tris[3 * (size_t)slot + 2] = make_float4(e2x, e2y, e2z, 0.f);
    }
    CUC(c->d_nodes.ensure(c->bvh.nodes.size() / 4));
    CUC(c->d_tris.ensure(h_tris.size()));
    CUC(cudaMemcpyAsync(c->d_nodes.p, c->bvh.nodes.data(), c->bvh.nodes.size() * sizeof(float), cudaMemcpyHostToDevice,
                        c->stream));
    CUC(cudaMemcpyAsync(c->d_tris.p, h_tris.data(), h_tris.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  }
  c->stats.bvh_build_ms = now_ms() - t_bvh0;
  if (std::max(c->bvh.depth, c->bvh.mesh_depth + kMaxRoots) > kStackDepth) {
    rt_destroy(c);
    return fail(RT_ERR_INVALID, "BVH deeper than the traversal stack");
  }
  c->stats.bvh_nodes = (int)c->bvh.num_nodes;
  c->stats.bvh_depth = c->bvh.depth;
  CUC(c->d_counters.ensure(kCntNum));
  CUC(cudaMemsetAsync(c->d_counters.p, 0, sizeof(unsigned long long) * kCntNum, c->stream));
  CUC(c->d_kd_pos.ensure(1));
  CUC(c->d_kd_dir.ensure(1));

  DScene& S = c->scene;
  S.nodes = c->d_nodes.p;
  S.tris = c->d_tris.p;
  S.tri_vidx = c->d_tri_vidx.p;
  S.pos = c->d_pos.p;
  S.nrm = c->d_nrm.p;
  S.mats = c->d_mats.p;
  S.num_lights = c->L;
  S.num_tris = c->T;
  {
    std::vector<DLight> ext;
    for (int l = 0; l < c->L; l++) {
      const rt_light& a = s->lights[l];
      const DLight d{h3(a.position), h3(a.color), h3(a.normal), h3(a.vertical), h3(a.horizontal),
                     a.intensity,    a.side,      a.ac,         a.al,           a.aq,
                     a.factor};
      if (l < kMaxLights)
        S.lights[l] = d;
      else
        ext.push_back(d);
    }
    S.lights_ext = nullptr;
    if (!ext.empty()) {  // Scene::lightsources() may hold any number of lights (Scene.h:14-26, Renderer.cpp:49)
      CUC(c->d_lights_ext.ensure(ext.size()));
      CUC(cudaMemcpyAsync(c->d_lights_ext.p, ext.data(), ext.size() * sizeof(DLight), cudaMemcpyHostToDevice, c->stream));
      S.lights_ext = c->d_lights_ext.p;
    }
  }
  S.cam = DCamera{h3(s->camera.position), h3(s->camera.lower_left), h3(s->camera.horizontal), h3(s->camera.vertical)};
  S.kd_pos = c->d_kd_pos.p;
  S.kd_dir = c->d_kd_dir.p;
  S.kd_count = 0;
  const int nroots = (int)(c->bvh.roots.size() / 7);
  S.num_roots = (nroots >= 2 && nroots <= kMaxRoots && !getenv("RT_NO_ROOT_LIST")) ? nroots : 0;
  for (int r = 0; r < S.num_roots; r++) {
    const float* q = c->bvh.roots.data() + 7 * r;
    S.root_lo[r] = make_float4(q[0], q[1], q[2], 0.f);
    S.root_hi[r] = make_float4(q[3], q[4], q[5], q[6]);
  }
  CUC(cudaStreamSynchronize(c->stream));  // uploads done: the host staging vectors above go out of scope
#undef CUC
  c->stats.create_ms = now_ms() - t_create0;
  *out = c;
  return RT_OK;
}

int rt_render_accumulate_device(rt_ctx* c, float* sum_rgb_device, int32_t* counter_device) {
  if (!c || !sum_rgb_device || !counter_device) return fail(RT_ERR_INVALID, "null argument");
  return render_to_device(c, sum_rgb_device, counter_device);
}

int rt_render_accumulate_packed_device(rt_ctx* c, float* sum_rgbn_device) {
  if (!c || !sum_rgbn_device) return fail(RT_ERR_INVALID, "null argument");
  if ((size_t)sum_rgbn_device % 16) return fail(RT_ERR_INVALID, "the packed frame must be 16-byte aligned");
  if (c->params.num_rays >= (1 << 24)) return fail(RT_ERR_INVALID, "packed counters are exact only below 2^24 samples");
  return render_to_device(c, nullptr, nullptr, nullptr, nullptr, reinterpret_cast<float4*>(sum_rgbn_device));
}

int rt_render_accumulate(rt_ctx* c, float* sum_rgb, int32_t* counter) {
  if (!c || !sum_rgb || !counter) return fail(RT_ERR_INVALID, "null argument");
  int rc = bind(c);
  if (rc) return rc;
  const size_t npx = (size_t)c->params.width * c->params.height;
  CU(c->d_out_rgb.ensure(3 * npx));
  CU(c->d_out_cnt.ensure(npx));
  if ((rc = render_to_device(c, c->d_out_rgb.p, c->d_out_cnt.p))) return rc;
  CU(cudaMemcpyAsync(sum_rgb, c->d_out_rgb.p, sizeof(float) * 3 * npx, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(counter, c->d_out_cnt.p, sizeof(int) * npx, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RT_OK;
}

// Renderer.cpp:262-265 after the last sample (i + 1 == N):
//   saveImage = updateImage / float(N) + image * (N - counter) / float(N)
int rt_composite(int32_t width, int32_t height, int32_t num_rays, const float* sum_rgb, const int32_t* counter,
                 float* rgb_inout) {
  if (!sum_rgb || !counter || !rgb_inout || width < 1 || height < 1) return fail(RT_ERR_INVALID, "bad argument");
  if (num_rays < 1) return fail(RT_ERR_INVALID, "num_rays must be >= 1 to composite");
  const float fn = (float)num_rays;
  for (size_t px = 0; px < (size_t)width * height; px++) {
    const float miss = (float)(num_rays - counter[px]);
    for (int ch = 0; ch < 3; ch++) {
      volatile float a = sum_rgb[3 * px + ch] / fn;
      volatile float b = rgb_inout[3 * px + ch] * miss;
      volatile float cc = b / fn;
      rgb_inout[3 * px + ch] = a + cc;
    }
  }
  return RT_OK;
}

// the same composite on the packed frame (sums + counter as one float4 per pixel) a single reduce produces
int rt_composite_packed_device(rt_ctx* c, int32_t num_rays, const float* sum_rgbn_device, float* rgb_inout) {
  if (!c || !sum_rgbn_device || !rgb_inout) return fail(RT_ERR_INVALID, "null argument");
  if (num_rays < 1) return fail(RT_ERR_INVALID, "num_rays must be >= 1 to composite");
  int rc = bind(c);
  if (rc) return rc;
  const size_t npx = (size_t)c->params.width * c->params.height;
  CU(c->d_out_rgb.ensure(3 * npx));
  CU(cudaMemcpyAsync(c->d_out_rgb.p, rgb_inout, sizeof(float) * 3 * npx, cudaMemcpyHostToDevice, c->stream));
  launch_composite_packed(reinterpret_cast<const float4*>(sum_rgbn_device), (long long)npx, num_rays, c->d_out_rgb.p,
                          c->stream);
  c->stats.kernel_launches++;
  CU(cudaMemcpyAsync(rgb_inout, c->d_out_rgb.p, sizeof(float) * 3 * npx, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RT_OK;
}

int rt_composite_device(rt_ctx* c, int32_t num_rays, const float* sum_rgb_device, const int32_t* counter_device,
                        float* rgb_inout) {
  if (!c || !sum_rgb_device || !counter_device || !rgb_inout) return fail(RT_ERR_INVALID, "null argument");
  if (num_rays < 1) return fail(RT_ERR_INVALID, "num_rays must be >= 1 to composite");
  int rc = bind(c);
  if (rc) return rc;
  const size_t npx = (size_t)c->params.width * c->params.height;
  CU(c->d_out_rgb.ensure(3 * npx));
  CU(cudaMemcpyAsync(c->d_out_rgb.p, rgb_inout, sizeof(float) * 3 * npx, cudaMemcpyHostToDevice, c->stream));
  launch_composite_frame(sum_rgb_device, counter_device, (long long)npx, num_rays, c->d_out_rgb.p, c->stream);
  c->stats.kernel_launches++;
  CU(cudaMemcpyAsync(rgb_inout, c->d_out_rgb.p, sizeof(float) * 3 * npx, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RT_OK;
}

int rt_render(rt_ctx* c, float* rgb_inout) { return rt_render_progressive(c, rgb_inout, 0, nullptr, nullptr); }

int rt_render_progressive(rt_ctx* c, float* rgb_inout, int32_t every, rt_progress_fn fn, void* user) {
  if (!c || !rgb_inout) return fail(RT_ERR_INVALID, "null argument");
  const rt_params& p = c->params;
  const size_t npx = (size_t)p.width * p.height;
  if (p.num_rays < 1) {
    // Renderer.cpp:208,219,271: the sample loop does not run and `image = saveImage`, a default-constructed
    // Image(w, h) whose Vec3f pixels are zero (Image.h:12-15, Vec3.h:21) -- -N 0 yields a black frame.
    std::memset(rgb_inout, 0, sizeof(float) * 3 * npx);
    return RT_OK;
  }
  int rc = bind(c);
  if (rc) return rc;
  // background up, the whole Renderer::render on the device (composite included), frame down
  CU(c->d_out_rgb.ensure(3 * npx));
  Progress progress{every, fn, user};
  if ((rc = render_to_device(c, nullptr, nullptr, c->d_out_rgb.p, &progress, nullptr, rgb_inout))) return rc;
  CU(cudaMemcpyAsync(rgb_inout, c->d_out_rgb.p, sizeof(float) * 3 * npx, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (fn && every > 0) fn(user, p.num_rays, p.num_rays, rgb_inout);  // the last pass: update.ppm == the result
  return RT_OK;
}

int rt_render_samples(rt_ctx* c, int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t s0, int32_t s1, float* rgb,
                      uint8_t* found) {
  if (!c || !rgb) return fail(RT_ERR_INVALID, "null argument");
  int rc = bind(c);
  if (rc) return rc;
  const rt_params& p = c->params;
  if (x0 < 0 || y0 < 0 || x1 > p.width || y1 > p.height || x0 >= x1 || y0 >= y1 || s0 < 0 || s0 >= s1)
    return fail(RT_ERR_INVALID, "bad window / sample range");
  if ((rc = prepare_photons(c))) return rc;
  const int ww = x1 - x0, wh = y1 - y0, ns = s1 - s0;
  std::vector<int> map((size_t)ww * wh);
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++) map[(size_t)(y - y0) * ww + (x - x0)] = y * p.width + x;
  const size_t paths = map.size() * (size_t)ns;
  if (paths > ((size_t)1 << 30)) return fail(RT_ERR_INVALID, "window too large");
  const bool photons = use_photon_map(c);
  if (paths * (size_t)std::max(shadow_lights(c, photons), 1) >= ((size_t)1 << 32))
    return fail(RT_ERR_INVALID, "window too large");
  DevBuf<int> d_map;
  CU(d_map.ensure(map.size()));
  CU(cudaMemcpyAsync(d_map.p, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if ((rc = prepare_batch_buffers(c, photons)) || (rc = ensure_work(c, paths, p.mode == 1, shadow_lights(c, photons)))) {
    d_map.release();
    return rc;
  }
  RenderArgs a;
  fill_args(c, a, d_map.p, (int)map.size(), photons);
  reset_call_timing(c);
  rc = run_batch(c, a, s0, ns);
  std::vector<float4> h(paths);
  cudaError_t e = cudaSuccess;
  if (rc == RT_OK) {
    launch_finalize_paths(c->d_col0.p, c->d_col1.p, c->d_col2.p, (long long)paths, p.mode == 1 ? 1 : 0, c->stream);
    e = cudaMemcpyAsync(h.data(), c->d_col0.p, paths * sizeof(float4), cudaMemcpyDeviceToHost, c->stream);
  }
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) e = e2;
  d_map.release();
  if (rc) return rc;
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, cudaGetErrorString(e));
  if ((rc = collect_spans(c))) return rc;
  c->stats.samples += paths;
  if ((rc = pull_counters(c))) return rc;
  for (size_t i = 0; i < paths; i++) {
    rgb[3 * i] = h[i].x;
    rgb[3 * i + 1] = h[i].y;
    rgb[3 * i + 2] = h[i].z;
    if (found) found[i] = h[i].w != 0.f;
  }
  return RT_OK;
}

static int trace_common(rt_ctx* c, const rt_ray* rays, int64_t n, rt_hit* hits, uint8_t* occluded, int32_t flags) {
  int rc = bind(c);
  if (rc) return rc;
  if (n < 0 || (n > 0 && !rays)) return fail(RT_ERR_INVALID, "bad ray batch");
  if (n >= ((int64_t)1 << 31)) return fail(RT_ERR_INVALID, "too many rays in one call");
  if (n == 0) return RT_OK;
  std::vector<float4> ho((size_t)n), hd((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    ho[i] = make_float4(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2], 0.f);
    hd[i] = make_float4(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2], 0.f);
  }
  DevBuf<float4> d_o, d_d, d_h;
  DevBuf<unsigned char> d_occ;
  int ret = RT_OK;
  cudaError_t e = d_o.ensure((size_t)n);
  if (e == cudaSuccess) e = d_d.ensure((size_t)n);
  if (e == cudaSuccess) e = d_h.ensure((size_t)n);
  if (e == cudaSuccess) e = d_occ.ensure((size_t)n);
  if (e == cudaSuccess) e = c->d_qcount.ensure(kQNum);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_o.p, ho.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_d.p, hd.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    const int depth = trace_stack_depth(c);
    const int persistent = c->num_sms * trace_ctas_per_sm(depth);
    int grid = (int)std::min<int64_t>((n + kBlock - 1) / kBlock, persistent);
    launch_trace_rays(c->scene, d_o.p, d_d.p, (unsigned)n, d_h.p, d_occ.p, occluded ? 1 : 0,
                      (flags & RT_FLAG_BRUTE_FORCE) ? 1 : 0, depth, c->d_qcount.p + kQFetchNearest0, std::max(grid, 1),
                      c->stream);
    c->stats.kernel_launches++;
    e = cudaStreamSynchronize(c->stream);
  }
  if (e == cudaSuccess) {
    if (occluded) {
      e = cudaMemcpy(occluded, d_occ.p, (size_t)n, cudaMemcpyDeviceToHost);
    } else {
      std::vector<float4> hh((size_t)n);
      e = cudaMemcpy(hh.data(), d_h.p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost);
      for (int64_t i = 0; i < n && e == cudaSuccess; i++) {
        int tri;
        std::memcpy(&tri, &hh[i].w, 4);
        hits[i] = tri >= 0 ? rt_hit{tri, hh[i].y, hh[i].z, hh[i].x} : rt_hit{-1, 0.f, 0.f, 0.f};
      }
    }
  }
  if (e != cudaSuccess) ret = fail(RT_ERR_CUDA, cudaGetErrorString(e));
  d_o.release();
  d_d.release();
  d_h.release();
  d_occ.release();
  return ret;
}

int rt_trace_rays(rt_ctx* c, const rt_ray* rays, int64_t n, rt_hit* hits, int32_t flags) {
  if (n > 0 && !hits) return fail(RT_ERR_INVALID, "null hits");
  return trace_common(c, rays, n, hits, nullptr, flags);
}
int rt_occluded(rt_ctx* c, const rt_ray* rays, int64_t n, uint8_t* occluded, int32_t flags) {
  if (n > 0 && !occluded) return fail(RT_ERR_INVALID, "null output");
  return trace_common(c, rays, n, nullptr, occluded, flags);
}

int rt_eval_bsdf(rt_ctx* c, const rt_material* m, const float* n_wi_wo, int64_t n, float* rgb) {
  int rc = bind(c);
  if (rc) return rc;
  if (!m || n < 0 || (n > 0 && (!n_wi_wo || !rgb))) return fail(RT_ERR_INVALID, "bad argument");
  if (n == 0) return RT_OK;
  DevBuf<float> d_in, d_out;
  cudaError_t e = d_in.ensure(9 * (size_t)n);
  if (e == cudaSuccess) e = d_out.ensure(3 * (size_t)n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in.p, n_wi_wo, sizeof(float) * 9 * (size_t)n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    launch_bsdf(make_material(m->kd, m->alpha, h3(m->albedo), h3(m->f0)), d_in.p, n, d_out.p, c->stream);
    c->stats.kernel_launches++;
    e = cudaStreamSynchronize(c->stream);
  }
  if (e == cudaSuccess) e = cudaMemcpy(rgb, d_out.p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost);
  d_in.release();
  d_out.release();
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, cudaGetErrorString(e));
  return RT_OK;
}

int rt_eval_hsphere(rt_ctx* c, uint64_t seed, uint64_t domain, uint64_t index0, const float* normals, int64_t n,
                    float* directions) {
  int rc = bind(c);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!normals || !directions))) return fail(RT_ERR_INVALID, "bad argument");
  if (n == 0) return RT_OK;
  DevBuf<float> d_in, d_out;
  cudaError_t e = d_in.ensure(3 * (size_t)n);
  if (e == cudaSuccess) e = d_out.ensure(3 * (size_t)n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in.p, normals, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    launch_hsphere(mix64(seed + kGolden), domain, index0, d_in.p, n, d_out.p, c->stream);
    c->stats.kernel_launches++;
    e = cudaMemcpyAsync(directions, d_out.p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  d_in.release();
  d_out.release();
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, cudaGetErrorString(e));
  return RT_OK;
}

int rt_photons_per_light(const rt_ctx* c, int32_t* out) {
  if (!c || !out) return fail(RT_ERR_INVALID, "null argument");
  *out = photons_per_light(c, nullptr);
  return RT_OK;
}

// Emission + ordered compaction on the device.  out7_dev: capacity particles of 7 floats (device memory of c's GPU).
static int emit_to_device(rt_ctx* c, int32_t first_path, int32_t num_paths, float* out7_dev, int64_t capacity,
                          int64_t* per_light_counts, int32_t* depth_histogram20, int64_t* stored,
                          bool count_only = false) {
  float light_pdf = 0.f;
  const int per_light = photons_per_light(c, &light_pdf);
  if (first_path < 0) first_path = 0;
  const int last = num_paths < 0 ? per_light : std::min(per_light, first_path + num_paths);
  const int npaths = std::max(0, last - first_path);
  if (per_light_counts)
    for (int l = 0; l < c->L; l++) per_light_counts[l] = 0;
  if (depth_histogram20)
    for (int i = 0; i < 20; i++) depth_histogram20[i] = 0;
  if (stored) *stored = 0;
  if (npaths == 0 || c->L == 0) return RT_OK;
  const size_t total = (size_t)c->L * npaths;
  if (total >= ((size_t)1 << 32)) return fail(RT_ERR_INVALID, "too many photon paths in one emission");
  const int nb = photon_compact_blocks((long long)total);
  DevBuf<float4> d_a, d_b;
  DevBuf<unsigned> d_blk, d_hist;
  DevBuf<unsigned long long> d_lc;
  std::vector<unsigned long long> h_lc(c->L);
  unsigned h_hist[20], h_total = 0;
  float ms = 0.f;
  cudaError_t e = d_a.ensure(total);
  if (e == cudaSuccess) e = d_b.ensure(total);
  if (e == cudaSuccess) e = d_blk.ensure((size_t)nb + 1);
  if (e == cudaSuccess) e = d_hist.ensure(21);  // 20 histogram bins + the emission kernel's work cursor
  if (e == cudaSuccess) e = d_lc.ensure((size_t)c->L);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_hist.p, 0, 20 * sizeof(unsigned), c->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_lc.p, 0, (size_t)c->L * sizeof(unsigned long long), c->stream);
  if (e == cudaSuccess) {
    cudaEventRecord(c->ev0, c->stream);
    launch_emit(c->scene, mix64(c->params.seed + kGolden), per_light, light_pdf, first_path, npaths,
                (c->params.flags & RT_FLAG_BRUTE_FORCE) ? 1 : 0, d_a.p, d_b.p, c->d_counters.p, d_hist.p + 20, c->num_sms,
                c->stream);
    cudaEventRecord(c->ev1, c->stream);
    launch_photon_compact(d_a.p, d_b.p, (long long)total, npaths, d_blk.p, d_lc.p, d_hist.p, out7_dev, capacity,
                          c->stream);
    c->stats.kernel_launches += 4;
    c->stats.kernel_count[kKEmit]++;
    e = cudaMemcpyAsync(h_lc.data(), d_lc.p, (size_t)c->L * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_hist, d_hist.p, sizeof(h_hist), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_total, d_blk.p + nb, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  }
  d_a.release();
  d_b.release();
  d_blk.release();
  d_hist.release();
  d_lc.release();
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, cudaGetErrorString(e));
  c->stats.photon_ms = ms;
  c->stats.kernel_ms[kKEmit] += ms;
  if (per_light_counts)
    for (int l = 0; l < c->L; l++) per_light_counts[l] = (int64_t)h_lc[l];
  if (depth_histogram20)
    for (int i = 0; i < 20; i++) depth_histogram20[i] = (int32_t)h_hist[i];
  if (stored) *stored = (int64_t)h_total;
  if (!count_only && (int64_t)h_total > capacity) return fail(RT_ERR_INVALID, "photon output capacity too small");
  return pull_counters(c);
}

int rt_emit_photons_device(rt_ctx* c, int32_t first_path, int32_t num_paths, float* out7_device, int64_t capacity,
                           int64_t* per_light_counts, int32_t* depth_histogram20) {
  int rc = bind(c);
  if (rc) return rc;
  if (!out7_device || capacity < 0) return fail(RT_ERR_INVALID, "bad device output");
  return emit_to_device(c, first_path, num_paths, out7_device, capacity, per_light_counts, depth_histogram20, nullptr);
}

int rt_emit_photons(rt_ctx* c, int32_t first_path, int32_t num_paths, rt_photon* out, int64_t capacity,
                    int64_t* per_light_counts, int32_t* depth_histogram20) {
  int rc = bind(c);
  if (rc) return rc;
  const int per_light = photons_per_light(c, nullptr);
  const int64_t dev_cap = out ? std::max<int64_t>(0, std::min<int64_t>(capacity, (int64_t)per_light * std::max(c->L, 0))) : 0;
  DevBuf<float> d_out;
  CU(d_out.ensure(7 * (size_t)std::max<int64_t>(dev_cap, 1)));
  int64_t stored = 0;
  // without an output buffer only the counts are wanted: the compaction writes nothing past capacity 0
  rc = emit_to_device(c, first_path, num_paths, d_out.p, dev_cap, per_light_counts, depth_histogram20, &stored, !out);
  if (rc == RT_OK && out && stored > 0) {
    cudaError_t e = cudaMemcpyAsync(out, d_out.p, sizeof(float) * 7 * (size_t)stored, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = fail(RT_ERR_CUDA, cudaGetErrorString(e));
  }
  d_out.release();
  return rc;
}

// The canonical tree of the exact k-NN mode (RT_FLAG_KNN_EXACT) is built on the device (csrc/kd_build.cu) unless
// RT_KD_BUILD=host; the reference-exact gather needs libstdc++'s nth_element tie placement and always builds on the host.
static bool kd_build_on_device(const rt_ctx* c) {
  if (!(c->params.flags & RT_FLAG_KNN_EXACT)) return false;
  const char* e = getenv("RT_KD_BUILD");
  return !(e && std::string(e) == "host");
}
static void kd_installed(rt_ctx* c, int64_t n) {
  c->scene.kd_pos = c->d_kd_pos.p;
  c->scene.kd_dir = c->d_kd_dir.p;
  c->scene.kd_count = (int)n;
  c->kd_n = n;
  c->stats.photons_stored = n;
  c->photon_map_built = true;
  c->photon_map_seed = c->params.seed;
  c->photon_map_requested = c->params.num_photons;
}
// the list (emission order, 7 floats per particle) is in device memory: build the canonical tree there
static int install_photons_device_build(rt_ctx* c, const float* d_p7, int64_t n) {
  const double t0 = now_ms();
  CU(c->d_kd_pos.ensure((size_t)std::max<int64_t>(n, 1)));
  CU(c->d_kd_dir.ensure((size_t)std::max<int64_t>(n, 1)));
  CU(c->d_kd_orig.ensure((size_t)std::max<int64_t>(n, 1)));
  std::string err;
  long long launches = 0;
  if (!build_kdtree_device(d_p7, (int)n, c->d_kd_pos.p, c->d_kd_dir.p, c->d_kd_orig.p, c->stream, &c->kd_height, &launches, err))
    return fail(RT_ERR_CUDA, "device kd-tree build failed: " + err);
  CU(cudaStreamSynchronize(c->stream));
  c->stats.kernel_launches += (uint64_t)launches;
  c->stats.kd_build_ms = now_ms() - t0;
  if (c->kd_height > kKdStack) return fail(RT_ERR_INVALID, "kd-tree deeper than the device stack");
  c->kd_on_device = true;
  c->kd_host_valid = false;  // rt_get_photons / rt_get_kdtree fetch the nodes on demand
  c->kd_nodes7.clear();
  kd_installed(c, n);
  return RT_OK;
}
// host copy of the kd-ordered nodes for rt_get_photons / rt_get_kdtree
static int ensure_kd_host(rt_ctx* c) {
  if (c->kd_host_valid) return RT_OK;
  const int64_t n = c->kd_n;
  std::vector<float4> hp((size_t)std::max<int64_t>(n, 1)), hd((size_t)std::max<int64_t>(n, 1));
  if (n > 0) {
    CU(cudaMemcpyAsync(hp.data(), c->d_kd_pos.p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hd.data(), c->d_kd_dir.p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  c->kd_nodes7.resize(7 * (size_t)n);
  for (int64_t i = 0; i < n; i++) {
    float* a = &c->kd_nodes7[7 * (size_t)i];
    a[0] = hp[i].x, a[1] = hp[i].y, a[2] = hp[i].z, a[3] = hd[i].x, a[4] = hd[i].y, a[5] = hd[i].z, a[6] = hp[i].w;
  }
  c->kd_host_valid = true;
  return RT_OK;
}

// shared tail of rt_set_photons / rt_set_photons_device: c->kd_nodes7 holds the list in emission order
static int install_photons(rt_ctx* c, int64_t n) {
  const double t_kd0 = now_ms();
  if (c->params.flags & RT_FLAG_KNN_EXACT)  // the canonical tree (RT_KD_BUILD=host; the device builder makes the same one)
    build_kdtree_canonical(c->kd_nodes7, &c->kd_height, nullptr);
  else
    build_kdtree(c->kd_nodes7, &c->kd_height);
  c->stats.kd_build_ms = now_ms() - t_kd0;
  if (c->kd_height > kKdStack) return fail(RT_ERR_INVALID, "kd-tree deeper than the device stack");
  DevBuf<float> d_p7;
  CU(d_p7.ensure(7 * (size_t)std::max<int64_t>(n, 1)));
  CU(c->d_kd_pos.ensure((size_t)std::max<int64_t>(n, 1)));
  CU(c->d_kd_dir.ensure((size_t)std::max<int64_t>(n, 1)));
  if (n > 0) {  // 28 bytes per particle up (not two padded float4 arrays), unpacked on the device
    CU(cudaMemcpyAsync(d_p7.p, c->kd_nodes7.data(), sizeof(float) * 7 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    launch_photon_unpack(d_p7.p, n, c->d_kd_pos.p, c->d_kd_dir.p, c->stream);
    c->stats.kernel_launches++;
  }
  CU(cudaStreamSynchronize(c->stream));
  d_p7.release();
  c->kd_on_device = false;
  c->kd_host_valid = true;
  kd_installed(c, n);
  return RT_OK;
}

int rt_set_photons(rt_ctx* c, const rt_photon* photons, int64_t n) {
  int rc = bind(c);
  if (rc) return rc;
  if (n < 0 || (n > 0 && !photons)) return fail(RT_ERR_INVALID, "bad photon list");
  if (n >= (1 << 28)) return fail(RT_ERR_INVALID, "too many photons");
  static_assert(sizeof(rt_photon) == 28, "rt_photon must match Particle (28 bytes)");
  if (kd_build_on_device(c)) {
    DevBuf<float> d_p7;
    CU(d_p7.ensure(7 * (size_t)std::max<int64_t>(n, 1)));
    if (n > 0) CU(cudaMemcpyAsync(d_p7.p, photons, sizeof(float) * 7 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    rc = install_photons_device_build(c, d_p7.p, n);
    d_p7.release();
    return rc;
  }
  c->kd_nodes7.assign((const float*)photons, (const float*)photons + 7 * n);
  return install_photons(c, n);
}

int rt_set_photons_device(rt_ctx* c, const float* photons7_device, int64_t n) {
  int rc = bind(c);
  if (rc) return rc;
  if (n < 0 || (n > 0 && !photons7_device)) return fail(RT_ERR_INVALID, "bad photon list");
  if (n >= (1 << 28)) return fail(RT_ERR_INVALID, "too many photons");
  if (kd_build_on_device(c)) return install_photons_device_build(c, photons7_device, n);  // nothing leaves the GPU
  // the kd-tree's shape is libstdc++'s nth_element: ONE device->host copy for the host build
  c->kd_nodes7.resize(7 * (size_t)n);
  if (n > 0) {
    CU(cudaMemcpyAsync(c->kd_nodes7.data(), photons7_device, sizeof(float) * 7 * (size_t)n, cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return install_photons(c, n);
}

int rt_splice_photons_device(rt_ctx* c, const float* gathered_device, int32_t world, int64_t stride,
                             const int64_t* counts, float* out7_device, int64_t capacity, int64_t* total_out) {
  int rc = bind(c);
  if (rc) return rc;
  if (!gathered_device || !counts || !out7_device || !total_out || world < 1 || stride < 0)
    return fail(RT_ERR_INVALID, "bad argument");
  const int L = c->L;
  std::vector<long long> seg_src, seg_dst;
  std::vector<long long> rank_off((size_t)world, 0);  // running offset inside every rank's shard
  long long total = 0;
  for (int l = 0; l < L; l++)
    for (int r = 0; r < world; r++) {
      const long long n = counts[(size_t)r * L + l];
      if (n < 0 || rank_off[r] + n > stride) return fail(RT_ERR_INVALID, "shard counts exceed the shard stride");
      if (n > 0) {
        seg_src.push_back((long long)r * stride + rank_off[r]);
        seg_dst.push_back(total);
      }
      rank_off[r] += n;
      total += n;
    }
  *total_out = total;
  if (total > capacity) return fail(RT_ERR_INVALID, "photon output capacity too small");
  if (total == 0) return RT_OK;
  seg_dst.push_back(total);
  const int nseg = (int)seg_src.size();
  DevBuf<long long> d_seg;
  CU(d_seg.ensure((size_t)2 * nseg + 1));
  CU(cudaMemcpyAsync(d_seg.p, seg_src.data(), sizeof(long long) * nseg, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_seg.p + nseg, seg_dst.data(), sizeof(long long) * (nseg + 1), cudaMemcpyHostToDevice, c->stream));
  launch_photon_splice(gathered_device, d_seg.p, d_seg.p + nseg, nseg, total, out7_device, c->stream);
  c->stats.kernel_launches++;
  cudaError_t e = cudaStreamSynchronize(c->stream);
  d_seg.release();
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, cudaGetErrorString(e));
  return RT_OK;
}

int rt_build_photon_map(rt_ctx* c) {
  int rc = bind(c);
  if (rc) return rc;
  // emission and ordered compaction on the device, one device->host copy of the stored particles for the kd build
  const int per_light = photons_per_light(c, nullptr);
  const int64_t cap = std::max<int64_t>(1, (int64_t)per_light * c->L);
  DevBuf<float> d_list;
  CU(d_list.ensure(7 * (size_t)cap));
  int64_t stored = 0;
  rc = emit_to_device(c, 0, -1, d_list.p, cap, nullptr, nullptr, &stored);
  if (rc == RT_OK) rc = rt_set_photons_device(c, d_list.p, stored);
  d_list.release();
  return rc;
}

int rt_get_photons(rt_ctx* c, rt_photon* out, int64_t capacity, int64_t* count) {
  if (!c || !count) return fail(RT_ERR_INVALID, "null argument");
  int rc0 = bind(c);
  if (rc0) return rc0;
  if ((rc0 = ensure_kd_host(c))) return rc0;
  int64_t n = (int64_t)c->kd_nodes7.size() / 7;
  *count = n;
  if (out) {
    if (capacity < n) return fail(RT_ERR_INVALID, "capacity too small");
    std::memcpy(out, c->kd_nodes7.data(), sizeof(float) * 7 * (size_t)n);
  }
  return RT_OK;
}

int rt_get_kdtree(rt_ctx* c, rt_photon* nodes, int32_t* left, int32_t* right, int32_t* root, int64_t capacity) {
  if (!c || !root) return fail(RT_ERR_INVALID, "null argument");
  int rc0 = bind(c);
  if (rc0) return rc0;
  if ((rc0 = ensure_kd_host(c))) return rc0;
  int64_t n = (int64_t)c->kd_nodes7.size() / 7;
  if (capacity < n) return fail(RT_ERR_INVALID, "capacity too small");
  if (nodes) std::memcpy(nodes, c->kd_nodes7.data(), sizeof(float) * 7 * (size_t)n);
  std::vector<int32_t> l, r;
  kdtree_links(n, l, r, root);
  if (left) std::memcpy(left, l.data(), sizeof(int32_t) * (size_t)n);
  if (right) std::memcpy(right, r.data(), sizeof(int32_t) * (size_t)n);
  return RT_OK;
}

int rt_knn(rt_ctx* c, const float* queries, int64_t n, int32_t k, int32_t* node_index) {
  int rc = bind(c);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!queries || !node_index))) return fail(RT_ERR_INVALID, "bad argument");
  if (c->scene.kd_count == 0) return fail(RT_ERR_EMPTY_TREE, "tree is empty");  // kdtree.h:181
  if (k > c->scene.kd_count) return fail(RT_ERR_K_TOO_LARGE, "k is greater than the number of nodes");
  if (k < 1 || k > RT_MAX_K) return fail(RT_ERR_INVALID, "k must be in [1, " + std::to_string(RT_MAX_K) + "]");
  if (n == 0) return RT_OK;
  // grid-stride kernel with exactly the resident CTA count (beyond kKnnSharedMaxK the candidates of every resident
  // thread live in a global scratch sized for that grid)
  const int ctas = (int)std::min<int64_t>((n + kBlock - 1) / kBlock,
                                          (int64_t)c->num_sms * knn_ctas_per_sm(k, c->kd_height + 1));
  if ((rc = ensure_knn_scratch(c, k, ctas * kBlock))) return rc;
  DevBuf<float> d_q;
  DevBuf<int> d_idx;
  cudaError_t e = d_q.ensure(3 * (size_t)n);
  if (e == cudaSuccess) e = d_idx.ensure((size_t)n * k);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_q.p, queries, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    launch_knn(c->scene, d_q.p, n, k, c->kd_height + 1, knn_flavour_k(c->params, k), d_idx.p, c->d_counters.p,
               c->d_knn_scratch.p, ctas, c->stream);
    c->stats.kernel_launches++;
    e = cudaStreamSynchronize(c->stream);
  }
  if (e == cudaSuccess) e = cudaMemcpy(node_index, d_idx.p, sizeof(int) * (size_t)n * k, cudaMemcpyDeviceToHost);
  d_q.release();
  d_idx.release();
  if (e != cudaSuccess) return fail(RT_ERR_CUDA, cudaGetErrorString(e));
  return RT_OK;
}

int rt_shard_pixels(const rt_params* p, int32_t* out, int64_t capacity, int64_t* count) {
  int rc = validate_params(p);
  if (rc) return rc;
  if (!count) return fail(RT_ERR_INVALID, "null count");
  std::vector<int> map;
  build_pix_map(*p, map);
  *count = (int64_t)map.size();
  if (out) {
    if (capacity < (int64_t)map.size()) return fail(RT_ERR_INVALID, "capacity too small");
    std::memcpy(out, map.data(), map.size() * sizeof(int));
  }
  return RT_OK;
}

int rt_profiler_range(int on) {
  CU(on ? cudaProfilerStart() : cudaProfilerStop());
  return RT_OK;
}

int rt_get_stats(rt_ctx* c, rt_stats* out) {
  if (!c || !out) return fail(RT_ERR_INVALID, "null argument");
  if (c->counters_dirty) {
    int rc = bind(c);
    if (rc) return rc;
    if ((rc = read_counters(c))) return rc;
  }
  *out = c->stats;
  return RT_OK;
}
int rt_reset_stats(rt_ctx* c) {
  if (!c) return fail(RT_ERR_INVALID, "null context");
  int nodes = c->stats.bvh_nodes, depth = c->stats.bvh_depth;
  int64_t stored = c->stats.photons_stored;
  const double cms = c->stats.create_ms, bms = c->stats.bvh_build_ms, kms = c->stats.kd_build_ms;
  int rc = bind(c);
  if (rc) return rc;
  CU(cudaMemsetAsync(c->d_counters.p, 0, sizeof(unsigned long long) * kCntNum, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->counters_dirty = false;
  c->stats = rt_stats{};
  c->stats.bvh_nodes = nodes;
  c->stats.bvh_depth = depth;
  c->stats.photons_stored = stored;
  c->stats.create_ms = cms;
  c->stats.bvh_build_ms = bms;
  c->stats.kd_build_ms = kms;
  return RT_OK;
}
int rt_build_bvh_host(const rt_scene* s, float pad_fraction, float* nodes16, int64_t capacity_nodes,
                      int32_t* slot_triangle, int32_t* num_nodes, int32_t* depth) {
  if (!s || s->num_triangles < 0 || s->num_meshes < 0) return fail(RT_ERR_INVALID, "bad scene");
  for (int t = 0; t < 3 * s->num_triangles; t++)
    if (s->triangles[t] < 0 || s->triangles[t] >= s->num_vertices)
      return fail(RT_ERR_INVALID, "triangle vertex index out of range");
  Bvh bvh;
  build_bvh(s->num_vertices, s->positions, s->num_triangles, s->triangles, s->num_meshes, s->mesh_first_triangle,
            scene_extent(s), pad_fraction > 0.f ? pad_fraction : 1.0f / 16384.0f, bvh);
  const int64_t n = (int64_t)bvh.nodes.size() / 16;
  if (num_nodes) *num_nodes = (int32_t)n;
  if (depth) *depth = bvh.depth;
  if (!nodes16) return RT_OK;
  if (capacity_nodes < n) return fail(RT_ERR_INVALID, "capacity too small");
  std::memcpy(nodes16, bvh.nodes.data(), bvh.nodes.size() * sizeof(float));
  if (slot_triangle) std::memcpy(slot_triangle, bvh.slot_tri.data(), bvh.slot_tri.size() * sizeof(int32_t));
  return RT_OK;
}
int rt_build_kdtree_host(float* photons7_inout, int64_t n, int32_t canonical, int32_t* orig_index, int32_t* height) {
  if (n < 0 || (n > 0 && !photons7_inout) || n >= (1 << 28)) return fail(RT_ERR_INVALID, "bad photon list");
  std::vector<float> v(photons7_inout, photons7_inout + 7 * n);
  std::vector<int32_t> orig;
  int h = 0;
  if (canonical)
    build_kdtree_canonical(v, &h, &orig);
  else
    build_kdtree(v, &h);
  std::memcpy(photons7_inout, v.data(), sizeof(float) * 7 * (size_t)n);
  if (orig_index && canonical) std::memcpy(orig_index, orig.data(), sizeof(int32_t) * (size_t)n);
  if (height) *height = h;
  return RT_OK;
}
int rt_get_bvh_slots(rt_ctx* c, int32_t* slot_triangle, int64_t capacity) {
  if (!c || !slot_triangle) return fail(RT_ERR_INVALID, "null argument");
  if (capacity < (int64_t)c->bvh.slot_tri.size()) return fail(RT_ERR_INVALID, "capacity too small");
  std::memcpy(slot_triangle, c->bvh.slot_tri.data(), c->bvh.slot_tri.size() * sizeof(int32_t));
  return RT_OK;
}
int rt_get_bvh(rt_ctx* c, float* nodes16, int64_t capacity_nodes, int32_t* num_nodes, int32_t* depth) {
  if (!c) return fail(RT_ERR_INVALID, "null context");
  int n = (int)c->bvh.num_nodes;
  if (num_nodes) *num_nodes = n;
  if (depth) *depth = c->bvh.depth;
  if (nodes16) {
    if (capacity_nodes < n) return fail(RT_ERR_INVALID, "capacity too small");
    if (c->bvh_on_device) {
      int rc = bind(c);
      if (rc) return rc;
      CU(cudaMemcpy(nodes16, c->d_nodes.p, c->bvh.num_nodes * 16 * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
      std::memcpy(nodes16, c->bvh.nodes.data(), c->bvh.nodes.size() * sizeof(float));
    }
  }
  return RT_OK;
}

}  // extern "C"
