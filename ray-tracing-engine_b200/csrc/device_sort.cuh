// device_sort.cuh -- pieces shared by the two level-synchronous device builders (bvh_build.cu, kd_build.cu): the
// order-preserving binary32 -> uint32 map and a bitonic sort of packed 64-bit words.  Everything has internal linkage
// (each translation unit gets its own copy of the kernels).
#pragma once
#include <cuda_runtime.h>

#include <cstring>

namespace rtb {
namespace {

__host__ __device__ inline unsigned f2ord(float f) {  // order-preserving map binary32 -> uint32
  unsigned u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  std::memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ inline float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// ---- bitonic sort of n = 2^m 64-bit words --------------------------------------------------------------------
constexpr int kSortTile = 2048;  // words per CTA in the shared-memory passes (1024 threads)
__device__ inline void cmp_swap(unsigned long long& x, unsigned long long& y, bool up) {
  if ((x > y) == up) {
    const unsigned long long t = x;
    x = y;
    y = t;
  }
}
// Every kernel sorts gridDim.y independent arrays of the same length in one launch: array y starts at a + y * stride.
// all steps (k, j) with k <= kSortTile: sorts every tile, alternating direction so that the merge can continue
__global__ void __launch_bounds__(1024) k_bitonic_tile_sort(unsigned long long* a, long long stride) {
  __shared__ unsigned long long s[kSortTile];
  a += (long long)blockIdx.y * stride;
  const long long base = (long long)blockIdx.x * kSortTile;
  s[threadIdx.x] = a[base + threadIdx.x];
  s[threadIdx.x + 1024] = a[base + threadIdx.x + 1024];
  __syncthreads();
  for (int k = 2; k <= kSortTile; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int i = 2 * threadIdx.x - (threadIdx.x & (j - 1));  // lower index of the pair
      const bool up = (((base + i) & k) == 0);
      cmp_swap(s[i], s[i + j], up);
      __syncthreads();
    }
  }
  a[base + threadIdx.x] = s[threadIdx.x];
  a[base + threadIdx.x + 1024] = s[threadIdx.x + 1024];
}
// one global step (k, j) with j >= kSortTile
__global__ void k_bitonic_global(unsigned long long* a, long long stride, long long n, long long k, long long j) {
  a += (long long)blockIdx.y * stride;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n / 2) return;
  const long long i = 2 * t - (t & (j - 1));
  const bool up = ((i & k) == 0);
  unsigned long long x = a[i], y = a[i + j];
  if ((x > y) == up) {
    a[i] = y;
    a[i + j] = x;
  }
}
// the steps j = kSortTile/2 .. 1 of merge size k > kSortTile, inside shared memory
__global__ void __launch_bounds__(1024) k_bitonic_tile_merge(unsigned long long* a, long long stride, long long k) {
  __shared__ unsigned long long s[kSortTile];
  a += (long long)blockIdx.y * stride;
  const long long base = (long long)blockIdx.x * kSortTile;
  s[threadIdx.x] = a[base + threadIdx.x];
  s[threadIdx.x + 1024] = a[base + threadIdx.x + 1024];
  __syncthreads();
  const bool up = ((base & k) == 0);  // the whole tile lies in one half of the k-block
  for (int j = kSortTile >> 1; j > 0; j >>= 1) {
    const int i = 2 * threadIdx.x - (threadIdx.x & (j - 1));
    cmp_swap(s[i], s[i + j], up);
    __syncthreads();
  }
  a[base + threadIdx.x] = s[threadIdx.x];
  a[base + threadIdx.x + 1024] = s[threadIdx.x + 1024];
}
// `batch` arrays of n2 words each, `stride` words apart, sorted by the same launches
bool sort_keys(unsigned long long* keys, long long n2, cudaStream_t st, long long* launches, int batch = 1,
               long long stride = 0) {
  k_bitonic_tile_sort<<<dim3((unsigned)(n2 / kSortTile), batch), 1024, 0, st>>>(keys, stride);
  ++*launches;
  for (long long k = 2 * kSortTile; k <= n2; k <<= 1) {
    for (long long j = k >> 1; j >= kSortTile; j >>= 1) {
      k_bitonic_global<<<dim3((unsigned)((n2 / 2 + 255) / 256), batch), 256, 0, st>>>(keys, stride, n2, k, j);
      ++*launches;
    }
    k_bitonic_tile_merge<<<dim3((unsigned)(n2 / kSortTile), batch), 1024, 0, st>>>(keys, stride, k);
    ++*launches;
  }
  return cudaGetLastError() == cudaSuccess;
}

}  // namespace
}  // namespace rtb
