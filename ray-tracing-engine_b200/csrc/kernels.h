// kernels.h -- launch interface between the C ABI (capi.cu) and the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "rt_device.cuh"

namespace rtb {

// Device counters (unsigned long long each)
enum Counter { kCntNearest = 0, kCntShadow = 1, kCntKnn = 2, kCntPhotonRays = 3, kCntKdVisits = 4, kCntNum = 8 };

// q_count[] slots (unsigned each): hits compacted by segment 0/1/2, then the work-fetch cursors of the
// persistent trace kernels (nearest-hit and any-hit launch of each segment)
enum QSlot { kQHits0 = 0, kQFetchNearest0 = 4, kQFetchAny0 = 8, kQNum = 12 };

// Shadow rays / light contributions of compacted hit j live in a light-major blocked layout so that a
// warp of the any-hit kernel gets 32 consecutive hit points aiming at ONE light (coherent rays):
//   slot(j, l) = (j / 32) * 32*nl + l * 32 + (j % 32)
// nl = shadow slots per hit: the scene's light count for direct lighting (Renderer.cpp:49 loops over
// scene.lightsources() of any length), 1 for the photon gather (one contribution, no shadow ray).
__host__ __device__ __forceinline__ unsigned shadow_slot(unsigned j, unsigned l, unsigned nl) {
  return (j >> 5) * (32u * nl) + l * 32u + (j & 31u);
}
__host__ __device__ __forceinline__ size_t shadow_slots_for(size_t hits, unsigned nl) {
  return ((hits + 31u) >> 5) * (size_t)(32u * nl);
}
// hit index j of shadow slot s (inverse of shadow_slot over l)
__host__ __device__ __forceinline__ unsigned shadow_slot_hit(unsigned s, unsigned nl) {
  if (nl == 3u) return (s / 96u) * 32u + (s & 31u);  // the stock three lights: a constant divisor (3 instructions, not ~20)
  return (s / (32u * nl)) * 32u + (s & 31u);
}

// Kernel classes timed separately with CUDA events on the launching stream (rt_stats::kernel_ms): bench.py
// picks the dominant one by its measured share of the step.
enum KernelClass {
  kKRaygen = 0, kKTraceNearest = 1, kKSort = 2, kKShade = 3, kKTraceAny = 4, kKCombine = 5, kKResolve = 6,
  kKEmit = 7, kKOther = 8, kKGather = 9, kKNumClasses = 10
};

// One wavefront batch = `nsamp` consecutive samples of `npix` pixels; path p = s_local*npix + pixel_local.
struct RenderArgs {
  DScene scene;
  int width, height;
  int num_rays;     // N of the whole render (stratification, RayTracer.h:111)
  int jitter_d;     // int(sqrt(float(N)))
  int mode;         // 0 ray, 1 path
  int photon;       // 1: gather from the photon map instead of direct lighting
  int k;            // neighbours
  int knn_exact;    // k-NN flavour: 1 = RT_FLAG_KNN_EXACT (canonical exact), 0 = reference search with the ascending
                    // candidate array (small k), -1 = reference search with libstdc++'s heap restated (large k)
  int kd_frames;    // kd-tree height + 1: stack frames per thread of kd_knearest_sorted
  int num_sms;      // multiprocessors of the device (grid sizing of the grid-stride kernels)
  int num_photons;  // REQUESTED photon count (Renderer.cpp:99)
  int brute;        // 1: O(T) scan instead of BVH
  int stack_depth;  // traversal stack entries per ray (>= bvh depth)
  uint64_t seed_mixed;
  const int* pix_map;  // local pixel -> y*W + x
  int npix;
  int s0, nsamp;
  float4* col0;  // per path: seg-0 colour (-m 0: clamped, final); w = posIntersectionFound
  float4* col1;  // per path: seg-1 colour (zero when the path ended before)
  float4* col2;  // per path: seg-2 colour (zero when the path ended before)
  float4* ray_o[2];  // ray queues, ping-pong by segment: (origin, path id) / (direction, -)
  float4* ray_d[2];
  float4* hit;       // per ray slot of the current segment: (t, u, v, triangle id | -1)
  int nl;            // shadow slots per hit (see shadow_slot): number of lights, or 1 for the photon gather
  int own_tri;       // 1: k_shade tests a shadow ray against the triangle it starts on before queueing it
  int sort_mask;     // bit s: segment s's hit points are binned by Morton cell before k_shade (default: the bounce
                     // segments 1 and 2; with a photon map also segment 0, whose queries otherwise keep the pixel-tile
                     // order of the primary rays)
  float4* hit_p;     // per compacted hit j: the hit point = origin of its nl shadow rays
  float4* sh_d;      // shadow ray directions, blocked layout (see shadow_slot); w != 0: already known occluded
  float4* contrib;   // radiance * bsdf of light l for hit j, same layout
  unsigned char* occ;  // any-hit result, same layout
  int* hit_path;     // compacted hit j -> path id
  // spatial binning of the hit points of a bounce segment (counting sort by Morton cell): perm[k] = ray slot
  // processed k-th by k_shade, so that consecutive hits (and the shadow rays they spawn) are neighbours
  unsigned int* perm;        // non-null: the bounce segments (and photon gathers) are processed in sorted order;
                             // holds the Morton cell of every ray between k_sort_count and k_sort_scatter
  float4* sorted;            // sorted payload, 32 B per ray: [2k] hit record, [2k+1] direction (xyz) + path id (w)
                             // of the ray processed k-th by k_shade
  unsigned int* sort_hist;   // kSortBuckets + 2 counters (bucket kSortBuckets = misses); with sort_bits > 0:
                             // kSortFineMax + 2 counters followed by kSortFineTiles tile sums of the scan
  int sort_bits;             // 0: 32^3 cells counted in shared memory; 6 / 7: (2^bits)^3 cells, histogram in global memory
  int tile_rounds;           // photon k_shade: > 1: order every tile of tile_rounds * kBlock slots by fine Morton code
                             // (host side only: negative = that many rounds whatever the amount of work, see launch_shade)
  float sort_key_scale;      // 1024 / cells per axis: sort_inv_cell * this maps a coordinate to the 10-bit fine grid
  float3 sort_lo;            // scene bounds
  float3 sort_inv_cell;      // cells per axis / extent per axis
  unsigned int* q_count;
  unsigned long long* counters;
  // photon gather by the persistent k_knn_gather (reference-exact flavours): per ray slot of the segment
  // (sum of the k incomeDirections in ascending distance, distance of the k-th); null: k_shade runs the query itself
  float4* knn_out;
  // k > kKnnSharedMaxK: the k-NN candidates of every resident thread live in global memory, [slot][thread]
  unsigned long long* knn_scratch;
  int knn_scratch_stride;  // threads the scratch was sized for (grid of the k-NN kernel * kBlock)
};
constexpr int kKnnSharedMaxK = 64;  // up to here the candidate rows fit in shared memory (2 CTAs/SM at 64)

#ifndef RT_SORT_GRID
#define RT_SORT_GRID 32
#endif
constexpr int kSortGrid = RT_SORT_GRID;                          // cells per axis (16: trace +0.3 ms, binning -0.3 ms: same frame time)
constexpr int kSortBuckets = kSortGrid * kSortGrid * kSortGrid;  // 32768 Morton cells (+1 bucket for misses)
constexpr int kSortFineMax = 1 << 21;                            // cells of the finest grid (128 per axis)
constexpr int kSortFineTiles = 1024;                             // >= tiles of 8 192 counters the fine scan can need
void launch_sort_hits(const RenderArgs& a, int seg, cudaStream_t st);  // fills a.perm for segment seg
void launch_raygen(const RenderArgs& a, cudaStream_t st);
void launch_trace_nearest(const RenderArgs& a, int seg, int grid_ctas, cudaStream_t st);
void launch_shade(const RenderArgs& a, int seg, int grid_ctas, cudaStream_t st);
void launch_trace_any(const RenderArgs& a, int seg, int grid_ctas, cudaStream_t st);
void launch_combine(const RenderArgs& a, int seg, int grid_ctas, cudaStream_t st);
int trace_ctas_per_sm(int stack_depth);      // resident CTAs per SM: nearest-hit kernel
int trace_any_ctas_per_sm(int stack_depth);  // ... any-hit kernel (smaller stack)
size_t trace_smem_bytes(int stack_depth);

void launch_resolve(const float4* col0, const float4* col1, const float4* col2, int mode, int npix, int nsamp,
                    float4* acc_rgb, int* acc_cnt, cudaStream_t st);
// col0 <- the final clamped path colours (what k_resolve accumulates), for rt_render_samples
void launch_finalize_paths(float4* col0, const float4* col1, const float4* col2, long long n, int mode, cudaStream_t st);
void launch_scatter(const float4* acc_rgb, const int* acc_cnt, const int* pix_map, int npix, float* out_rgb,
                    int* out_cnt, cudaStream_t st);
// sums + counter as one float4 per pixel (one reduce across GPUs instead of two)
void launch_scatter_packed(const float4* acc_rgb, const int* acc_cnt, const int* pix_map, int npix, float4* out,
                           cudaStream_t st);
void launch_composite_packed(const float4* sum_rgbn, long long npx, int num_rays, float* rgb_inout, cudaStream_t st);
// the final composite over the background on the device (rt_render): rgb_inout holds the background on entry
void launch_composite_frame(const float* sum_rgb, const int* counter, long long npx, int num_rays, float* rgb_inout,
                            cudaStream_t st);
void launch_composite(const float4* acc_rgb, const int* acc_cnt, const int* pix_map, int npix, int num_rays,
                      const float* background, float* out, cudaStream_t st);
// parity hooks: caller-supplied rays through the SAME persistent trace kernels the renderer uses
void launch_trace_rays(const DScene& s, const float4* ro, const float4* rd, unsigned n, float4* hits,
                       unsigned char* occluded, int any, int brute, int stack_depth, unsigned* fetch_counter,
                       int grid_ctas, cudaStream_t st);
size_t knn_smem_bytes(int k, int kd_frames);                   // dynamic shared memory of the kernels that run the k-NN
int shade_photon_ctas_per_sm(int mode, int k, int kd_frames);  // resident CTAs of the k-NN shade kernel
// persistent k-nearest-photon gather of a segment's hit points (fills a.knn_out); grid = SMs x knn_gather_ctas_per_sm
void launch_knn_gather(const RenderArgs& a, int seg, cudaStream_t st);
int knn_gather_ctas_per_sm(int k, int kd_frames, int flavour);
void launch_bsdf(DMaterial m, const float* n_wi_wo, long long n, float* rgb, cudaStream_t st);
void launch_hsphere(uint64_t seed_mixed, uint64_t domain, uint64_t index0, const float* normals, long long n, float* out,
                    cudaStream_t st);
// photon emission: path q = light*npaths + j traces path (first_path + j) of `light`
// cursor: one unsigned in device memory (zeroed by the launch), the work-fetch cursor of the persistent lanes
void launch_emit(const DScene& s, uint64_t seed_mixed, int per_light, float light_pdf, int first_path, int npaths,
                 int brute, float4* out_a, float4* out_b, unsigned long long* counters, unsigned* cursor, int num_sms,
                 cudaStream_t st);
// device-resident photon list: ordered compaction of k_emit's output, splice of all-gathered shards, unpack for the gather
int photon_compact_blocks(long long total);
void launch_photon_compact(const float4* out_a, const float4* out_b, long long total, int npaths, unsigned* block_count,
                           unsigned long long* light_count, unsigned* hist20, float* out7, long long capacity,
                           cudaStream_t st);
void launch_photon_splice(const float* gathered, const long long* seg_src, const long long* seg_dst, int nseg,
                          long long total, float* out7, cudaStream_t st);
void launch_photon_unpack(const float* p7, long long n, float4* kd_pos, float4* kd_dir, cudaStream_t st);
// scratch: knn candidates in global memory when k > kKnnSharedMaxK (k * 8 bytes per thread of `grid_ctas` CTAs)
void launch_knn(const DScene& s, const float* q3, long long n, int k, int kd_frames, int exact, int* node_index,
                unsigned long long* counters, unsigned long long* scratch, int grid_ctas, cudaStream_t st);
int knn_ctas_per_sm(int k, int kd_frames);

}  // namespace rtb
