// kd_build.cu -- the photon kd-tree of kdtree::make_tree (source/kdtree.h:60-69) as a level-synchronous GPU build
// (SURVEY.md 8f-2), for the CANONICAL tree of the exact k-NN mode (RT_FLAG_KNN_EXACT).
//
// kdtree::make_tree puts the median of [begin, end) on the cycling axis at n = begin + (end - begin)/2 with
// std::nth_element, then recurses into [begin, n) and [n+1, end): the array ends in in-order layout and the child
// links are implied by the ranges.  Which photon becomes the median when coordinates tie (20-29 % of the photons
// share an exactly equal coordinate: they sit on axis-aligned walls) is whatever libstdc++'s introselect leaves --
// the reference-exact gather therefore keeps the host build (csrc/host_build.cpp).  The exact mode does not depend
// on that accident: its tree orders the photons of a range by (coordinate, index in the emitted list), a TOTAL order,
// so the tree is a function of the photon list alone, the host builder (build_kdtree_canonical) and this one produce
// the identical array, and the median split becomes order statistics that every level can do at once:
//
//   1. three global orders: photons sorted by (coordinate_a, index) for a = x, y, z (bitonic sort of one packed
//      64-bit word per photon);
//   2. per level (axis = depth mod 3), for all ranges of the level together:
//        k_kd_mark       the photon at the median position of its range in the level axis' order becomes the node at
//                        that array position; every other photon of the range gets a side (left / right) by its rank
//        k_kd_flags      per position of the two OTHER orders: packed (goes left, goes right) flags
//        scan            exclusive scan of the packed flags (64-bit: both counts at once)
//        k_kd_partition  stable partition of the two other orders inside every range into [left | median | right];
//                        the level axis' own order already is partitioned
//      A range occupies the same positions [b, e) in all three orders, so the only per-position state is (b, e).
//   3. nothing else: the node array IS the result (positions + directions in kd order).
//
// Depth = floor(log2 n) + 1 levels, ~6 launches each; 357 835 photons (cfg 4) build in a couple of ms against 9 ms
// for the multi-threaded host nth_element, and the list never leaves the GPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <utility>

#include "device_sort.cuh"
#include "kd_build.h"

namespace rtb {
namespace {

#define KD(call)                                                     \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      err = std::string(#call) + ": " + cudaGetErrorString(e_);      \
      return false;                                                  \
    }                                                                \
  } while (0)

__global__ void k_kd_keys(const float* __restrict__ p7, int n, unsigned long long* kx, unsigned long long* ky,
                          unsigned long long* kz, int2* seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* a = p7 + 7 * (size_t)i;
  kx[i] = ((unsigned long long)f2ord(a[0]) << 32) | (unsigned)i;
  ky[i] = ((unsigned long long)f2ord(a[1]) << 32) | (unsigned)i;
  kz[i] = ((unsigned long long)f2ord(a[2]) << 32) | (unsigned)i;
  seg[i] = make_int2(0, n);  // level 0: one range, the whole array
}
__global__ void k_kd_fill(unsigned long long* p, long long from, long long to) {
  const long long i = from + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < to) p[i] = ~0ull;
}
__global__ void k_kd_extract(const unsigned long long* __restrict__ keys, int n, int* ord) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ord[i] = (int)(unsigned)keys[i];
}

// side: 0 left, 1 the median (this level's node), 2 right; seg_next: the range of the position on the next level
__global__ void k_kd_mark(const int* __restrict__ ord_axis, const int2* __restrict__ seg, int n,
                          const float* __restrict__ p7, unsigned char* side, int2* seg_next, float4* kd_pos,
                          float4* kd_dir, int* kd_orig) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int2 s = seg[i];
  if (s.x < 0) {  // a node placed on an earlier level
    seg_next[i] = s;
    return;
  }
  const int m = s.x + (s.y - s.x) / 2;  // kdtree.h:62
  const int p = ord_axis[i];
  if (i == m) {
    const float* a = p7 + 7 * (size_t)p;
    kd_pos[m] = make_float4(a[0], a[1], a[2], a[6]);  // Particle{position, incomeDirection, weight}
    kd_dir[m] = make_float4(a[3], a[4], a[5], 0.f);
    kd_orig[m] = p;
    side[p] = 1;
    seg_next[i] = make_int2(-1, -1);
  } else if (i < m) {
    side[p] = 0;
    seg_next[i] = make_int2(s.x, m);
  } else {
    side[p] = 2;
    seg_next[i] = make_int2(m + 1, s.y);
  }
}
struct KdOrders {
  const int* in[2];
  int* out[2];
};
// packed flags of the two other orders (blockIdx.y): low word 1 = goes left, high word 1 = goes right
__global__ void k_kd_flags(KdOrders o, const int2* __restrict__ seg, const unsigned char* __restrict__ side, int n,
                           unsigned long long* flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long f = 0;
  if (seg[i].x >= 0) {
    const unsigned char c = side[o.in[blockIdx.y][i]];
    f = c == 0 ? 1ull : (c == 2 ? (1ull << 32) : 0ull);
  }
  flags[(size_t)blockIdx.y * n + i] = f;
}
constexpr int kKdScanBlock = 1024;
__global__ void __launch_bounds__(kKdScanBlock) k_kd_scan_block(const unsigned long long* __restrict__ in_all, int n,
                                                                unsigned long long* out_all,
                                                                unsigned long long* block_sum_all) {
  __shared__ unsigned long long s[kKdScanBlock];
  const unsigned long long* in = in_all + (size_t)blockIdx.y * n;
  unsigned long long* out = out_all + (size_t)blockIdx.y * n;
  const int i = blockIdx.x * kKdScanBlock + threadIdx.x;
  const unsigned long long v = i < n ? in[i] : 0ull;
  s[threadIdx.x] = v;
  __syncthreads();
  for (int off = 1; off < kKdScanBlock; off <<= 1) {
    const unsigned long long add = threadIdx.x >= off ? s[threadIdx.x - off] : 0ull;
    __syncthreads();
    s[threadIdx.x] += add;
    __syncthreads();
  }
  if (i < n) out[i] = s[threadIdx.x] - v;
  if (threadIdx.x == kKdScanBlock - 1) block_sum_all[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s[threadIdx.x];
}
__global__ void __launch_bounds__(kKdScanBlock) k_kd_scan_sums(unsigned long long* block_sum_all, int nb) {
  __shared__ unsigned long long s[kKdScanBlock];
  unsigned long long* block_sum = block_sum_all + (size_t)blockIdx.x * nb;
  unsigned long long carry = 0;
  for (int base = 0; base < nb; base += kKdScanBlock) {
    const int i = base + threadIdx.x;
    const unsigned long long v = i < nb ? block_sum[i] : 0ull;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kKdScanBlock; off <<= 1) {
      const unsigned long long add = threadIdx.x >= off ? s[threadIdx.x - off] : 0ull;
      __syncthreads();
      s[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < nb) block_sum[i] = carry + s[threadIdx.x] - v;
    const unsigned long long total = s[kKdScanBlock - 1];
    __syncthreads();
    carry += total;
  }
}
__global__ void k_kd_partition(KdOrders o, const int2* __restrict__ seg, const unsigned char* __restrict__ side,
                               const unsigned long long* __restrict__ scan_all,
                               const unsigned long long* __restrict__ block_sum_all, int n, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = blockIdx.y;
  const int p = o.in[a][i];
  const int2 s = seg[i];
  int dst = i;
  if (s.x >= 0) {
    const unsigned long long* scan = scan_all + (size_t)a * n;
    const unsigned long long* bs = block_sum_all + (size_t)a * nb;
    const unsigned long long at_i = scan[i] + bs[i / kKdScanBlock], at_b = scan[s.x] + bs[s.x / kKdScanBlock];
    const unsigned long long d = at_i - at_b;  // (lefts, rights) of this range before position i: no borrow, both
    const int m = s.x + (s.y - s.x) / 2;       // counts are monotone
    const unsigned char c = side[p];
    dst = c == 0 ? s.x + (int)(unsigned)d : (c == 1 ? m : m + 1 + (int)(d >> 32));
  }
  o.out[a][dst] = p;
}

struct KdArena {
  char* base = nullptr;
  size_t used = 0, cap = 0;
  cudaStream_t st;
  explicit KdArena(cudaStream_t s) : st(s) {}
  static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
  template <typename T>
  T* take(size_t n) {
    T* p = reinterpret_cast<T*>(base + used);
    used += padded(std::max<size_t>(n, 1) * sizeof(T));
    return p;
  }
  ~KdArena() {
    if (base) cudaFreeAsync(base, st);
  }
};

}  // namespace

int kd_canonical_height(long long n) {
  int h = 0;
  while (n > 0) {  // the left child [b, n) holds floor(size / 2) photons, never fewer than the right one
    h++;
    n /= 2;
  }
  return h;
}

bool build_kdtree_device(const float* d_photons7, int n, float4* d_kd_pos, float4* d_kd_dir, int* d_kd_orig,
                         cudaStream_t st, int* height_out, long long* launches_out, std::string& err) {
  long long launches = 0;
  if (height_out) *height_out = kd_canonical_height(n);
  if (n < 1) return true;
  if (n >= (1 << 28)) {
    err = "too many photons for the device kd-tree builder";
    return false;
  }
  const int threads = 256, blocks = (n + threads - 1) / threads;
  const int nb = (n + kKdScanBlock - 1) / kKdScanBlock;
  long long n2 = kSortTile;
  while (n2 < n) n2 <<= 1;
  KdArena arena(st);
  {
    auto P = KdArena::padded;
    const size_t t = (size_t)n;
    arena.cap = 3 * P((size_t)n2 * 8) + 5 * P(t * 4) + 2 * P(t * 8) + P(t) + 2 * P(2 * t * 8) + P(2 * (size_t)nb * 8) + 4096;
    KD(cudaMallocAsync((void**)&arena.base, arena.cap, st));
  }
  unsigned long long* key[3] = {arena.take<unsigned long long>(n2), arena.take<unsigned long long>(n2),
                                arena.take<unsigned long long>(n2)};
  int* ord[3] = {arena.take<int>(n), arena.take<int>(n), arena.take<int>(n)};
  int* tmp[2] = {arena.take<int>(n), arena.take<int>(n)};
  int2* seg = arena.take<int2>(n);
  int2* seg_next = arena.take<int2>(n);
  unsigned char* side = arena.take<unsigned char>(n);
  unsigned long long* flags = arena.take<unsigned long long>(2 * (size_t)n);
  unsigned long long* scan = arena.take<unsigned long long>(2 * (size_t)n);
  unsigned long long* block_sum = arena.take<unsigned long long>(2 * (size_t)nb);
  if (arena.used > arena.cap) {
    err = "internal: kd scratch arena too small";
    return false;
  }
  k_kd_keys<<<blocks, threads, 0, st>>>(d_photons7, n, key[0], key[1], key[2], seg);
  launches++;
  for (int a = 0; a < 3 && n2 > n; a++) {
    k_kd_fill<<<(unsigned)((n2 - n + 255) / 256), 256, 0, st>>>(key[a], n, n2);
    launches++;
  }
  // the three key arrays are consecutive pieces of the arena (n2 * 8 bytes is a multiple of its 256-byte granule): one
  // batched sort, a third of the launches
  if (key[1] != key[0] + n2 || key[2] != key[1] + n2 || !sort_keys(key[0], n2, st, &launches, 3, n2)) {
    err = "bitonic sort launch failed";
    return false;
  }
  for (int a = 0; a < 3; a++) {
    k_kd_extract<<<blocks, threads, 0, st>>>(key[a], n, ord[a]);
    launches++;
  }
  const int levels = kd_canonical_height(n);
  for (int level = 0; level < levels; level++) {
    const int ax = level % 3;  // kdtree.h:66: index = (index + 1) % 3
    k_kd_mark<<<blocks, threads, 0, st>>>(ord[ax], seg, n, d_photons7, side, seg_next, d_kd_pos, d_kd_dir, d_kd_orig);
    launches++;
    if (level + 1 < levels) {
      const int o0 = (ax + 1) % 3, o1 = (ax + 2) % 3;
      KdOrders oo;
      oo.in[0] = ord[o0];
      oo.in[1] = ord[o1];
      oo.out[0] = tmp[0];
      oo.out[1] = tmp[1];
      k_kd_flags<<<dim3(blocks, 2), threads, 0, st>>>(oo, seg, side, n, flags);
      k_kd_scan_block<<<dim3(nb, 2), kKdScanBlock, 0, st>>>(flags, n, scan, block_sum);
      k_kd_scan_sums<<<2, kKdScanBlock, 0, st>>>(block_sum, nb);
      k_kd_partition<<<dim3(blocks, 2), threads, 0, st>>>(oo, seg, side, scan, block_sum, n, nb);
      launches += 4;
      std::swap(ord[o0], tmp[0]);
      std::swap(ord[o1], tmp[1]);
    }
    std::swap(seg, seg_next);
  }
  KD(cudaGetLastError());
  if (launches_out) *launches_out = launches;
  return true;
}

}  // namespace rtb
