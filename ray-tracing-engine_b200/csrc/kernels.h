// kernels.h -- launch interface between the C ABI (capi.cu) and the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "rt_device.cuh"

namespace rtb {

// Device counters (unsigned long long each)
enum Counter { kCntNearest = 0, kCntShadow = 1, kCntKnn = 2, kCntPhotonRays = 3, kCntKdVisits = 4, kCntNum = 8 };

// One wavefront batch = `nsamp` consecutive samples of `npix` pixels; path p = s_local*npix + pixel_local.
struct RenderArgs {
  DScene scene;
  int width, height;
  int num_rays;     // N of the whole render (stratification, RayTracer.h:111)
  int jitter_d;     // int(sqrt(float(N)))
  int mode;         // 0 ray, 1 path
  int photon;       // 1: gather from the photon map instead of direct lighting
  int k;            // neighbours
  int num_photons;  // REQUESTED photon count (Renderer.cpp:99)
  int brute;        // 1: O(T) scan instead of BVH
  uint64_t seed_mixed;
  const int* pix_map;  // local pixel -> y*W + x
  int npix;
  int s0, nsamp;
  float4* col0;  // per path: seg-0 colour, then the final clamped colour; w = posIntersectionFound
  float4* col1;  // per path: seg-1 colour
  float4* q_o[2];
  float4* q_d[2];           // ray queues (origin+path id, direction)
  unsigned int* q_count;    // [3] entries pushed by segment 0,1,2
  unsigned long long* counters;
};

void launch_segment(const RenderArgs& a, int seg, int grid_ctas, cudaStream_t st);
void launch_resolve(const float4* col0, int npix, int nsamp, float4* acc_rgb, int* acc_cnt, cudaStream_t st);
void launch_scatter(const float4* acc_rgb, const int* acc_cnt, const int* pix_map, int npix, float* out_rgb,
                    int* out_cnt, cudaStream_t st);
void launch_trace_rays(const DScene& s, const float* rays6, long long n, int* tri, float* uvt, int brute, int any,
                       unsigned char* occluded, cudaStream_t st);
void launch_bsdf(DMaterial m, const float* n_wi_wo, long long n, float* rgb, cudaStream_t st);
// photon emission: path q = light*npaths + j traces path (first_path + j) of `light`
void launch_emit(const DScene& s, uint64_t seed_mixed, int per_light, float light_pdf, int first_path, int npaths,
                 int brute, float4* out_a, float4* out_b, unsigned long long* counters, cudaStream_t st);
void launch_knn(const DScene& s, const float* q3, long long n, int k, int* node_index, unsigned long long* counters,
                cudaStream_t st);
int segment_ctas_per_sm(int mode, int photon);

}  // namespace rtb
