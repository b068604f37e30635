// bvh_build.cu -- the reference's BVH split policy (source/BVH.h:100-161) as a level-synchronous GPU build.
//
// the host build is O(T log T) but serial per subtree; at 1.2 M triangles it is the start-up
// bottleneck (0.3-0.6 s against 0.12 s of kernels per 32-sample pass on 8 GPUs).  The policy itself is simple --
//   node box = union of the node's triangles; leaf iff one triangle (:123); cut axis = first axis with the strictly
//   largest extent (:131-140); order the triangles by the key sum_v vertex[v][axis] (:141-150); left = the first
//   floor(n/2) (:151-158)
// -- and every level of the tree can be processed at once:
//
//   1. three global orders: triangles sorted by (mesh, key_a, index) for a = x, y, z (bitonic sort of one packed
//      64-bit word per triangle; the index makes the order total, so the build is deterministic);
//   2. per level, for all segments (= nodes) of that level together:
//        k_seg_box     box of every segment (ordered-int atomic min/max, warp-aggregated)
//        k_seg_split   axis, child ranges, child node indices (analytic: a subtree over n triangles owns n-1
//                      consecutive nodes, left child at node+1, right child at node+floor(n/2)), child boxes into
//                      the parent node, leaves into the slot table
//        k_mark        side of every triangle = its rank in the cut axis' order >= floor(n/2)
//        per axis: flags -> exclusive scan -> stable partition of that axis' order inside every segment
//   3. the leaf-order triangle array (p0, e1, e2, id) the traversal reads.
//
// A segment occupies the same position range [b, e) in all three orders, which are also the leaf slots of its
// subtree, so no per-triangle bookkeeping other than `side` is needed.  Among equal keys the reference's order is
// whatever libstdc++'s unstable std::sort leaves; ours is by triangle index.  The tree shape, the node numbering and
// the layout are those of csrc/host_build.cpp, which stays the builder for small scenes and the checker in the tests.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "bvh_build.h"
#include "device_sort.cuh"

namespace rtb {
namespace {

#define BV(call)                                                     \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      err = std::string(#call) + ": " + cudaGetErrorString(e_);      \
      return false;                                                  \
    }                                                                \
  } while (0)

struct Seg {  // one node of the current level
  int b, e;        // position range in the three orders == leaf slots of the subtree
  int node;        // index of this subtree's root node when e - b > 1
  int parent;      // node that holds this segment's box as child `side`; -1: a mesh root
  int side;
  int axis, child; // filled by k_seg_split: cut axis, index of the left child in the next level's table
};

// ---- per triangle: exact box and the three keys -------------------------------------------------------------
__global__ void k_tri_prepare(const float4* __restrict__ pos, const int4* __restrict__ vidx, int T, float4* box_lo,
                              float4* box_hi, unsigned long long* key0, unsigned long long* key1,
                              unsigned long long* key2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int4 v = vidx[t];
  const float4 a = pos[v.x], b = pos[v.y], c = pos[v.z];
  box_lo[t] = make_float4(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z), 0.f);
  box_hi[t] = make_float4(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z), 0.f);
  // BVH.h:143-148: vertex 0 + vertex 1 + vertex 2 on the axis, binary32, left to right
  const float kx = __fadd_rn(__fadd_rn(a.x, b.x), c.x), ky = __fadd_rn(__fadd_rn(a.y, b.y), c.y),
              kz = __fadd_rn(__fadd_rn(a.z, b.z), c.z);
  const unsigned long long hi = (unsigned long long)(unsigned)v.w << 56, lo = (unsigned)t;  // mesh | key | index
  key0[t] = hi | ((unsigned long long)f2ord(kx) << 24) | lo;
  key1[t] = hi | ((unsigned long long)f2ord(ky) << 24) | lo;
  key2[t] = hi | ((unsigned long long)f2ord(kz) << 24) | lo;
}
__global__ void k_fill_u64(unsigned long long* p, long long from, long long to, unsigned long long v) {
  const long long i = from + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < to) p[i] = v;
}

__global__ void k_extract_index(const unsigned long long* __restrict__ keys, int T, int* ord) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) ord[i] = (int)(keys[i] & 0xffffffull);
}

// ---- per level --------------------------------------------------------------------------------------------------
__global__ void k_seg_box(const int* __restrict__ seg_of_pos, const int* __restrict__ ord0,
                          const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, int T, unsigned* seg_box) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = i < T ? seg_of_pos[i] : -1;
  unsigned v[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  if (s >= 0) {
    const int t = ord0[i];
    const float4 lo = box_lo[t], hi = box_hi[t];
    v[0] = f2ord(lo.x), v[1] = f2ord(lo.y), v[2] = f2ord(lo.z);
    v[3] = f2ord(hi.x), v[4] = f2ord(hi.y), v[5] = f2ord(hi.z);
  }
  // warp aggregation: when the whole warp sits in one segment (the common case on the upper levels) one lane
  // issues the six atomics
  const unsigned full = 0xffffffffu;
  const int s0 = __shfl_sync(full, s, 0);
  if (__all_sync(full, s == s0)) {
    if (s0 < 0) return;
#pragma unroll
    for (int c = 0; c < 6; c++)
      for (int off = 16; off > 0; off >>= 1) {
        const unsigned o = __shfl_xor_sync(full, v[c], off);
        v[c] = c < 3 ? min(v[c], o) : max(v[c], o);
      }
    if ((threadIdx.x & 31) == 0) {
      for (int c = 0; c < 3; c++) atomicMin(seg_box + 6 * (size_t)s0 + c, v[c]);
      for (int c = 3; c < 6; c++) atomicMax(seg_box + 6 * (size_t)s0 + c, v[c]);
    }
  } else if (s >= 0) {
    for (int c = 0; c < 3; c++) atomicMin(seg_box + 6 * (size_t)s + c, v[c]);
    for (int c = 3; c < 6; c++) atomicMax(seg_box + 6 * (size_t)s + c, v[c]);
  }
}

__global__ void k_seg_split(Seg* seg, int S, const unsigned* __restrict__ seg_box, const int* __restrict__ ord0, float pad,
                            float* nodes, int* slot_tri, float* root_box, int* root_ref, Seg* next, int* next_count,
                            unsigned* next_box) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  Seg g = seg[s];
  const int n = g.e - g.b;
  float lo[3], hi[3];
  for (int c = 0; c < 3; c++) {
    lo[c] = ord2f(seg_box[6 * (size_t)s + c]);
    hi[c] = ord2f(seg_box[6 * (size_t)s + 3 + c]);
  }
  const int ref = n == 1 ? ~g.b : g.node;
  if (g.parent >= 0) {  // this segment is child `side` of its parent node (layout: csrc/host_build.h)
    float* nd = nodes + 16 * (size_t)g.parent + 6 * g.side;
    for (int c = 0; c < 3; c++) {
      nd[c] = lo[c] - pad;
      nd[3 + c] = hi[c] + pad;
    }
    nodes[16 * (size_t)g.parent + 12 + g.side] = __int_as_float(ref);
  } else {  // a mesh root: the host joins these (exact boxes)
    const int r = g.side;  // root slot
    for (int c = 0; c < 3; c++) {
      root_box[6 * r + c] = lo[c];
      root_box[6 * r + 3 + c] = hi[c];
    }
    root_ref[r] = ref;
  }
  if (n == 1) {
    slot_tri[g.b] = ord0[g.b];
    g.axis = -1;
    g.child = -1;
  } else {
    // BVH.h:131-140: the first axis whose extent is strictly larger than everything before it
    float longest = 0.f;
    int axis = 0;
    for (int c = 0; c < 3; c++) {
      const float len = hi[c] - lo[c];
      if (len > longest) {
        longest = len;
        axis = c;
      }
    }
    const int nl = n / 2;  // BVH.h:151-158
    const int c0 = atomicAdd(next_count, 2);
    next[c0] = Seg{g.b, g.b + nl, g.node + 1, g.node, 0, 0, 0};
    next[c0 + 1] = Seg{g.b + nl, g.e, g.node + nl, g.node, 1, 0, 0};
    for (int c = 0; c < 3; c++) {
      next_box[6 * (size_t)c0 + c] = next_box[6 * (size_t)(c0 + 1) + c] = 0xffffffffu;
      next_box[6 * (size_t)c0 + 3 + c] = next_box[6 * (size_t)(c0 + 1) + 3 + c] = 0u;
    }
    nodes[16 * (size_t)g.node + 14] = nodes[16 * (size_t)g.node + 15] = 0.f;
    g.axis = axis;
    g.child = c0;
  }
  seg[s] = g;
}

__global__ void k_mark(const Seg* __restrict__ seg, const int* __restrict__ seg_of_pos, const int* __restrict__ ord_x,
                       const int* __restrict__ ord_y, const int* __restrict__ ord_z, int T, unsigned char* side,
                       int* seg_next) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const int s = seg_of_pos[i];
  if (s < 0) {
    seg_next[i] = -1;
    return;
  }
  const Seg g = seg[s];
  if (g.axis < 0) {
    seg_next[i] = -1;  // a leaf: this position is final
    return;
  }
  const int nl = (g.e - g.b) / 2;
  const bool right = i >= g.b + nl;
  seg_next[i] = g.child + (right ? 1 : 0);
  const int* ord = g.axis == 0 ? ord_x : (g.axis == 1 ? ord_y : ord_z);
  side[ord[i]] = right ? 1 : 0;
}

// The three axis orders are processed by one launch each step (blockIdx.y = axis); per-axis arrays are strided by T
// (flags, scan, out) or by nb (block sums).
struct Orders {
  const int* in[3];
  int* out[3];
};
// flags of one axis' order: 1 = the triangle at this position moves to the right child of a segment cut on ANOTHER axis
__global__ void k_flags(const Seg* __restrict__ seg, const int* __restrict__ seg_of_pos, Orders ord,
                        const unsigned char* __restrict__ side, int T, int* flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int axis = blockIdx.y;
  if (i >= T) return;
  const int s = seg_of_pos[i];
  int f = 0;
  if (s >= 0) {
    const int a = seg[s].axis;
    if (a >= 0 && a != axis) f = side[ord.in[axis][i]];
  }
  flags[(size_t)axis * T + i] = f;
}
// exclusive scan of T ints: per-block scan + block sums, scan of the sums by one block, add back
constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(kScanBlock) k_scan_block(const int* __restrict__ in_all, int T, int* out_all,
                                                           int* block_sum_all) {
  __shared__ int s[kScanBlock];
  const int* in = in_all + (size_t)blockIdx.y * T;
  int* out = out_all + (size_t)blockIdx.y * T;
  int* block_sum = block_sum_all + (size_t)blockIdx.y * gridDim.x;
  const int i = blockIdx.x * kScanBlock + threadIdx.x;
  const int v = i < T ? in[i] : 0;
  s[threadIdx.x] = v;
  __syncthreads();
  for (int off = 1; off < kScanBlock; off <<= 1) {
    const int add = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
    __syncthreads();
    s[threadIdx.x] += add;
    __syncthreads();
  }
  if (i < T) out[i] = s[threadIdx.x] - v;
  if (threadIdx.x == kScanBlock - 1) block_sum[blockIdx.x] = s[threadIdx.x];
}
__global__ void __launch_bounds__(kScanBlock) k_scan_sums(int* block_sum_all, int nb) {  // one CTA per axis
  __shared__ int s[kScanBlock];
  int* block_sum = block_sum_all + (size_t)blockIdx.x * nb;
  int carry = 0;
  for (int base = 0; base < nb; base += kScanBlock) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? block_sum[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kScanBlock; off <<= 1) {
      const int add = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
      __syncthreads();
      s[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < nb) block_sum[i] = carry + s[threadIdx.x] - v;
    const int total = s[kScanBlock - 1];
    __syncthreads();
    carry += total;
  }
}
__global__ void k_partition(const Seg* __restrict__ seg, const int* __restrict__ seg_of_pos, Orders ord,
                            const int* __restrict__ flags_all, const int* __restrict__ scan_all,
                            const int* __restrict__ block_sum_all, int T, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int axis = blockIdx.y;
  if (i >= T) return;
  const int* flags = flags_all + (size_t)axis * T;
  const int* scan = scan_all + (size_t)axis * T;
  const int* block_sum = block_sum_all + (size_t)axis * nb;
  const int s = seg_of_pos[i];
  int dst = i;
  if (s >= 0) {
    const Seg g = seg[s];
    if (g.axis >= 0 && g.axis != axis) {
      const int nl = (g.e - g.b) / 2;
      const int fi = scan[i] + block_sum[i / kScanBlock], fb = scan[g.b] + block_sum[g.b / kScanBlock];
      const int r = fi - fb;  // triangles of this segment before position i that go right
      dst = flags[i] ? g.b + nl + r : g.b + (i - g.b) - r;
    }
  }
  ord.out[axis][dst] = ord.in[axis][i];
}
__global__ void k_init_pos(const Seg* __restrict__ seg, int S, int* seg_of_pos) {  // level 0: mesh ranges
  const int s = blockIdx.y;
  const Seg g = seg[s];
  for (int i = g.b + blockIdx.x * blockDim.x + threadIdx.x; i < g.e; i += gridDim.x * blockDim.x) seg_of_pos[i] = s;
}
// leaf-order triangles: (p0, global id) (e1) (e2) with the reference's own binary32 subtractions (Ray.cpp:11)
__global__ void k_leaf_tris(const float4* __restrict__ pos, const int4* __restrict__ vidx, const int* __restrict__ slot_tri,
                            int T, float4* tris) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= T) return;
  const int gid = slot_tri[slot];
  const int4 v = vidx[gid];
  const float4 a = pos[v.x], b = pos[v.y], c = pos[v.z];
  tris[3 * (size_t)slot] = make_float4(a.x, a.y, a.z, __int_as_float(gid));
  tris[3 * (size_t)slot + 1] = make_float4(__fsub_rn(b.x, a.x), __fsub_rn(b.y, a.y), __fsub_rn(b.z, a.z), 0.f);
  tris[3 * (size_t)slot + 2] = make_float4(__fsub_rn(c.x, a.x), __fsub_rn(c.y, a.y), __fsub_rn(c.z, a.z), 0.f);
}

// ---- small meshes: one thread-block cluster builds one mesh's whole subtree ---------------------------------------
// The level loop above costs 7 launches per level (120 launches, 0.64 ms at 11 666 triangles: launch latency, the
// kernels are empty).  When no mesh has more than kSmallMeshMax triangles, a cluster of kMeshCluster CTAs x 1024 threads
// runs the same phases for its mesh with the hardware cluster barrier between them; the scratch arrays stay in global
// memory and are read with ld.cg (L2: the CTAs of a cluster sit on different SMs, whose L1s are not coherent).
// A position's segment at a level -- its range [b, e), node index, parent and side -- follows from halving the mesh's
// range `level` times (node numbering is analytic, see above), so no segment table is kept: descend() recomputes it.
// (Measured, B200, 11 666 triangles in 5 meshes: a single CTA per mesh took 1.69 ms -- 32 warps walking 12 rounds of
// dependent L2 round trips per phase; see profiles/r2_tuning.md for the cluster's number.)
constexpr int kMeshThreads = 1024;
constexpr int kMeshCluster = 8;
constexpr int kSmallMeshMax = 16384;
struct MeshJob {
  int t0, n;      // the mesh's triangle range == its position range in the three orders == its leaf slots
  int node_base;  // first node of its subtree
  int root;       // slot in the root tables
};
struct SegAt {
  int b, e, node, parent, side;
  bool active;  // false: the position's segment became a leaf on an earlier level
};
__device__ inline SegAt descend(const MeshJob& J, int i, int level) {
  SegAt g{J.t0, J.t0 + J.n, J.node_base, -1, J.root, true};
  for (int l = 0; l < level; l++) {
    const int n = g.e - g.b;
    if (n < 2) {
      g.active = false;
      break;
    }
    const int nl = n / 2;  // BVH.h:151-158
    g.parent = g.node;
    if (i < g.b + nl) {
      g.side = 0;
      g.node = g.node + 1;
      g.e = g.b + nl;
    } else {
      g.side = 1;
      g.node = g.node + nl;
      g.b = g.b + nl;
    }
  }
  return g;
}
struct MeshScratch {
  const float4 *box_lo, *box_hi;     // per triangle
  const unsigned long long* key[3];  // sorted (mesh | key | index) words
  int *ord_a, *ord_b;                // 3 x T each: the three orders, ping-pong
  unsigned* seg_box;                 // 6 per position, used at a segment's first position
  int* seg_axis;                     // per position, used at a segment's first position
  unsigned char* side;               // per triangle
  int* scan;                         // 3 x T
  int* slot_tri;
  float* root_box;
  int* root_ref;
};
__global__ void __launch_bounds__(kMeshThreads) k_build_mesh(const MeshJob* __restrict__ jobs, MeshScratch W, int T, int levels_max,
                                                             const float4* __restrict__ pos, const int4* __restrict__ vidx,
                                                             float pad, float* nodes, float4* tris) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ unsigned long long warp_total[kMeshThreads / 32];
  __shared__ unsigned long long cta_total;
  const MeshJob J = jobs[blockIdx.x / kMeshCluster];
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, gt = rank * kMeshThreads + tid, GT = kMeshCluster * kMeshThreads;
  const int t0 = J.t0, t1 = J.t0 + J.n;
  int levels = 1;
  for (int n = J.n; n > 1; n -= n / 2) levels++;  // tree_depth(n)
  levels = min(levels, levels_max);
  int* ord_in = W.ord_a;
  int* ord_out = W.ord_b;
  for (int a = 0; a < 3; a++)
    for (int i = t0 + gt; i < t1; i += GT) ord_in[(size_t)a * T + i] = (int)(W.key[a][i] & 0xffffffull);
  if (gt < 6) W.seg_box[6 * (size_t)t0 + gt] = gt < 3 ? 0xffffffffu : 0u;
  cluster.sync();
  const int rounds = (J.n + GT - 1) / GT;  // positions per thread, strided (a warp = 32 neighbours)
  const int chunk = rounds;                // positions per thread, contiguous (the scans)
  for (int level = 0; level < levels; level++) {
    // A. segment boxes: ordered-int atomic min/max at the segment's first position, warp-aggregated
    for (int r = 0; r < rounds; r++) {
      const int i = t0 + r * GT + gt;
      int s = -1;
      unsigned v[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
      if (i < t1) {
        const SegAt g = descend(J, i, level);
        if (g.active) {
          s = g.b;
          const int t = __ldcg(ord_in + i);
          const float4 lo = W.box_lo[t], hi = W.box_hi[t];
          v[0] = f2ord(lo.x), v[1] = f2ord(lo.y), v[2] = f2ord(lo.z);
          v[3] = f2ord(hi.x), v[4] = f2ord(hi.y), v[5] = f2ord(hi.z);
        }
      }
      const unsigned full = 0xffffffffu;
      const int s0 = __shfl_sync(full, s, 0);
      if (__all_sync(full, s == s0)) {
        if (s0 >= 0) {
#pragma unroll
          for (int c = 0; c < 6; c++)
            for (int off = 16; off > 0; off >>= 1) {
              const unsigned o = __shfl_xor_sync(full, v[c], off);
              v[c] = c < 3 ? min(v[c], o) : max(v[c], o);
            }
          if ((tid & 31) == 0) {
            for (int c = 0; c < 3; c++) atomicMin(W.seg_box + 6 * (size_t)s0 + c, v[c]);
            for (int c = 3; c < 6; c++) atomicMax(W.seg_box + 6 * (size_t)s0 + c, v[c]);
          }
        }
      } else if (s >= 0) {
        for (int c = 0; c < 3; c++) atomicMin(W.seg_box + 6 * (size_t)s + c, v[c]);
        for (int c = 3; c < 6; c++) atomicMax(W.seg_box + 6 * (size_t)s + c, v[c]);
      }
    }
    cluster.sync();
    // B. one thread per segment (the one at its first position): the box goes into the parent node, a leaf into the
    //    slot table; an inner node picks its axis and clears its children's accumulators
    for (int r = 0; r < rounds; r++) {
      const int i = t0 + r * GT + gt;
      if (i >= t1) break;
      const SegAt g = descend(J, i, level);
      if (!g.active || i != g.b) continue;
      const int n = g.e - g.b;
      float lo[3], hi[3];
      for (int c = 0; c < 3; c++) {
        lo[c] = ord2f(__ldcg(W.seg_box + 6 * (size_t)g.b + c));
        hi[c] = ord2f(__ldcg(W.seg_box + 6 * (size_t)g.b + 3 + c));
      }
      const int ref = n == 1 ? ~g.b : g.node;
      if (g.parent >= 0) {
        float* nd = nodes + 16 * (size_t)g.parent + 6 * g.side;
        for (int c = 0; c < 3; c++) {
          nd[c] = lo[c] - pad;
          nd[3 + c] = hi[c] + pad;
        }
        nodes[16 * (size_t)g.parent + 12 + g.side] = __int_as_float(ref);
      } else {
        for (int c = 0; c < 3; c++) {
          W.root_box[6 * g.side + c] = lo[c];
          W.root_box[6 * g.side + 3 + c] = hi[c];
        }
        W.root_ref[g.side] = ref;
      }
      if (n == 1) {
        W.slot_tri[g.b] = __ldcg(ord_in + g.b);
      } else {
        float longest = 0.f;  // BVH.h:131-140: the first axis whose extent is strictly larger than everything before it
        int axis = 0;
        for (int c = 0; c < 3; c++) {
          const float len = hi[c] - lo[c];
          if (len > longest) {
            longest = len;
            axis = c;
          }
        }
        W.seg_axis[g.b] = axis;
        nodes[16 * (size_t)g.node + 14] = nodes[16 * (size_t)g.node + 15] = 0.f;
        const int nl = n / 2;
        for (int c = 0; c < 6; c++)
          W.seg_box[6 * (size_t)g.b + c] = W.seg_box[6 * (size_t)(g.b + nl) + c] = c < 3 ? 0xffffffffu : 0u;
      }
    }
    if (level + 1 >= levels) break;  // every segment of the last level is a leaf
    cluster.sync();
    // C. side of every triangle = its rank in the cut axis' order >= floor(n/2)
    for (int r = 0; r < rounds; r++) {
      const int i = t0 + r * GT + gt;
      if (i >= t1) break;
      const SegAt g = descend(J, i, level);
      if (!g.active || g.e - g.b < 2) continue;
      const int axis = __ldcg(W.seg_axis + g.b);
      W.side[__ldcg(ord_in + (size_t)axis * T + i)] = i >= g.b + (g.e - g.b) / 2 ? 1 : 0;
    }
    cluster.sync();
    // D1. per order: exclusive scan over the mesh's positions of "this triangle moves right" (0 inside segments cut on
    //     this very axis, inside leaves and finished segments).  A thread owns `chunk` consecutive positions; the three
    //     running sums travel as 21-bit fields of one word (n <= 16 384).
    {
      const int c0 = t0 + gt * chunk, c1 = min(c0 + chunk, t1);
      unsigned long long local = 0;
      for (int i = c0; i < c1; i++) {
        const SegAt g = descend(J, i, level);
        if (!g.active || g.e - g.b < 2) continue;
        const int axis = __ldcg(W.seg_axis + g.b);
        for (int a = 0; a < 3; a++)
          if (a != axis) local += (unsigned long long)__ldcg(W.side + __ldcg(ord_in + (size_t)a * T + i)) << (21 * a);
      }
      unsigned long long incl = local;
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, off);
        if ((tid & 31) >= off) incl += o;
      }
      if ((tid & 31) == 31) warp_total[tid >> 5] = incl;
      __syncthreads();
      if (tid < 32) {
        unsigned long long w = warp_total[tid], wi = w;
        for (int off = 1; off < 32; off <<= 1) {
          const unsigned long long o = __shfl_up_sync(0xffffffffu, wi, off);
          if (tid >= off) wi += o;
        }
        warp_total[tid] = wi - w;  // exclusive
        if (tid == 31) cta_total = wi;
      }
      cluster.sync();  // every CTA's total is in its shared memory (a __syncthreads as well)
      unsigned long long run = warp_total[tid >> 5] + incl - local;
      for (int q = 0; q < rank; q++) run += *cluster.map_shared_rank(&cta_total, q);  // distributed shared memory
      for (int i = c0; i < c1; i++) {
        for (int a = 0; a < 3; a++) W.scan[(size_t)a * T + i] = (int)((run >> (21 * a)) & 0x1fffffull);
        const SegAt g = descend(J, i, level);
        if (!g.active || g.e - g.b < 2) continue;
        const int axis = __ldcg(W.seg_axis + g.b);
        for (int a = 0; a < 3; a++)
          if (a != axis) run += (unsigned long long)__ldcg(W.side + __ldcg(ord_in + (size_t)a * T + i)) << (21 * a);
      }
    }
    cluster.sync();
    // D2. stable partition of the two other orders inside every segment
    for (int r = 0; r < rounds; r++) {
      const int i = t0 + r * GT + gt;
      if (i >= t1) break;
      const SegAt g = descend(J, i, level);
      const bool split = g.active && g.e - g.b >= 2;
      const int axis = split ? __ldcg(W.seg_axis + g.b) : -1;
      for (int a = 0; a < 3; a++) {
        const int t = __ldcg(ord_in + (size_t)a * T + i);
        int dst = i;
        if (split && a != axis) {
          const int nl = (g.e - g.b) / 2;
          // of this segment, before i, going right
          const int rr = __ldcg(W.scan + (size_t)a * T + i) - __ldcg(W.scan + (size_t)a * T + g.b);
          dst = __ldcg(W.side + t) ? g.b + nl + rr : g.b + (i - g.b) - rr;
        }
        ord_out[(size_t)a * T + dst] = t;
      }
    }
    cluster.sync();
    int* sw = ord_in;
    ord_in = ord_out;
    ord_out = sw;
  }
  cluster.sync();
  // leaf-order triangles: (p0, global id) (e1) (e2) with the reference's own binary32 subtractions (Ray.cpp:11)
  for (int slot = t0 + gt; slot < t1; slot += GT) {
    const int gid = __ldcg(W.slot_tri + slot);
    const int4 v = vidx[gid];
    const float4 a = pos[v.x], b = pos[v.y], c = pos[v.z];
    tris[3 * (size_t)slot] = make_float4(a.x, a.y, a.z, __int_as_float(gid));
    tris[3 * (size_t)slot + 1] = make_float4(__fsub_rn(b.x, a.x), __fsub_rn(b.y, a.y), __fsub_rn(b.z, a.z), 0.f);
    tris[3 * (size_t)slot + 2] = make_float4(__fsub_rn(c.x, a.x), __fsub_rn(c.y, a.y), __fsub_rn(c.z, a.z), 0.f);
  }
}

// one stream-ordered scratch allocation carved into 256-byte aligned pieces (25 separate cudaMallocAsync calls
// cost 9.5 ms of a 17 ms build)
struct Arena {
  char* base = nullptr;
  size_t used = 0, cap = 0;
  cudaStream_t st;
  explicit Arena(cudaStream_t s) : st(s) {}
  static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
  cudaError_t reserve(size_t bytes) {
    cap = bytes;
    return cudaMallocAsync((void**)&base, std::max<size_t>(bytes, 256), st);
  }
  template <typename T>
  T* take(size_t n) {
    T* p = reinterpret_cast<T*>(base + used);
    used += padded(std::max<size_t>(n, 1) * sizeof(T));
    return p;
  }
  ~Arena() {
    if (base) cudaFreeAsync(base, st);
  }
};

int tree_depth(int n) {  // nodes on the longest root-to-leaf path of the median-split tree over n triangles
  int d = 1;
  while (n > 1) {
    n -= n / 2;
    d++;
  }
  return d;
}

// the join of the per-mesh roots on the host, shared by both device paths (st is synchronised on return)
bool finish_build(int T, int nonempty, int max_n, float pad, float* d_root_box, int* d_root_ref, int* d_slot_tri,
                  float4* d_nodes, cudaStream_t st, Bvh& out, std::string& err) {
  std::vector<float> h_root_box(6 * (size_t)nonempty);
  std::vector<int32_t> h_root_ref(nonempty);
  out.slot_tri.resize(T);
  BV(cudaMemcpyAsync(h_root_box.data(), d_root_box, sizeof(float) * h_root_box.size(), cudaMemcpyDeviceToHost, st));
  BV(cudaMemcpyAsync(h_root_ref.data(), d_root_ref, sizeof(int32_t) * nonempty, cudaMemcpyDeviceToHost, st));
  BV(cudaMemcpyAsync(out.slot_tri.data(), d_slot_tri, sizeof(int32_t) * (size_t)T, cudaMemcpyDeviceToHost, st));
  BV(cudaStreamSynchronize(st));
  BV(cudaGetLastError());
  const size_t num_nodes = (size_t)(T - nonempty) + (size_t)std::max(nonempty - 1, 0);
  out.nodes.assign(16 * std::max<size_t>((size_t)std::max(nonempty - 1, 1), 1), 0.f);  // host copy of the join only
  build_top_level(nonempty, h_root_box.data(), h_root_ref.data(), pad, tree_depth(max_n), out);
  if (nonempty > 1)
    BV(cudaMemcpyAsync(d_nodes, out.nodes.data(), sizeof(float) * 16 * (size_t)(nonempty - 1), cudaMemcpyHostToDevice, st));
  BV(cudaStreamSynchronize(st));
  out.num_nodes = num_nodes;  // the caller fetches the nodes from the device on demand (rt_get_bvh)
  return true;
}

// keys + one batched sort + one cluster per mesh (k_build_mesh): 15 launches at 11 666 triangles instead of 120
bool build_small_meshes(const float4* d_pos, const int4* d_vidx, int T, const std::vector<Seg>& roots, int max_n,
                        long long n2, float pad, float4* d_nodes, float4* d_tris, cudaStream_t st, Bvh& out,
                        long long* launches_out, std::string& err) {
  long long launches = 0;
  const int nonempty = (int)roots.size();
  const int threads = 256, blocksT = (T + threads - 1) / threads;
  Arena arena(st);
  auto P = Arena::padded;
  const size_t t = (size_t)T;
  BV(arena.reserve(2 * P(t * 16) + 3 * P((size_t)n2 * 8) + 2 * P(3 * t * 4) + P(6 * t * 4) + P(t * 4) + P(t) + P(3 * t * 4) +
                   P(t * 4) + P(6 * (size_t)nonempty * 4) + P((size_t)nonempty * 4) + P((size_t)nonempty * sizeof(MeshJob)) +
                   4096));
  float4* box_lo = arena.take<float4>(T);
  float4* box_hi = arena.take<float4>(T);
  unsigned long long* key[3] = {arena.take<unsigned long long>(n2), arena.take<unsigned long long>(n2),
                                arena.take<unsigned long long>(n2)};
  MeshScratch W;
  W.box_lo = box_lo;
  W.box_hi = box_hi;
  for (int a = 0; a < 3; a++) W.key[a] = key[a];
  W.ord_a = arena.take<int>(3 * t);
  W.ord_b = arena.take<int>(3 * t);
  W.seg_box = arena.take<unsigned>(6 * t);
  W.seg_axis = arena.take<int>(T);
  W.side = arena.take<unsigned char>(T);
  W.scan = arena.take<int>(3 * t);
  W.slot_tri = arena.take<int>(T);
  W.root_box = arena.take<float>(6 * (size_t)nonempty);
  W.root_ref = arena.take<int>(nonempty);
  MeshJob* d_jobs = arena.take<MeshJob>(nonempty);
  if (arena.used > arena.cap || key[1] != key[0] + n2 || key[2] != key[1] + n2) {
    err = "internal: scratch arena too small";
    return false;
  }
  std::vector<MeshJob> jobs;  // stays alive until the synchronisation in finish_build
  for (const Seg& r : roots) jobs.push_back(MeshJob{r.b, r.e - r.b, r.node, r.side});
  BV(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(MeshJob) * jobs.size(), cudaMemcpyHostToDevice, st));
  k_tri_prepare<<<blocksT, threads, 0, st>>>(d_pos, d_vidx, T, box_lo, box_hi, key[0], key[1], key[2]);
  launches++;
  for (int a = 0; a < 3 && n2 > T; a++) {
    k_fill_u64<<<(unsigned)((n2 - T + 255) / 256), 256, 0, st>>>(key[a], T, n2, ~0ull);
    launches++;
  }
  if (!sort_keys(key[0], n2, st, &launches, 3, n2)) {
    err = "bitonic sort launch failed";
    return false;
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nonempty * kMeshCluster);
    cfg.blockDim = dim3(kMeshThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kMeshCluster;
    attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    const MeshJob* jobs_arg = d_jobs;
    BV(cudaLaunchKernelEx(&cfg, k_build_mesh, jobs_arg, W, T, tree_depth(max_n), d_pos, d_vidx, pad, (float*)d_nodes, d_tris));
    launches++;
  }
  BV(cudaGetLastError());
  if (!finish_build(T, nonempty, max_n, pad, W.root_box, W.root_ref, W.slot_tri, d_nodes, st, out, err)) return false;
  if (launches_out) *launches_out = launches;
  return true;
}

}  // namespace

bool build_bvh_device(const float4* d_pos, const int4* d_vidx, int T, int M, const int32_t* mesh_first_triangle,
                      float pad, float4* d_nodes, float4* d_tris, cudaStream_t st, Bvh& out, long long* launches_out,
                      std::string& err) {
  long long launches = 0;
  const bool timing = getenv("RT_BVH_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto tick = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[bvh_build] %s %.2f ms; ", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  std::vector<Seg> roots;
  int node_base = 0, max_n = 0;
  for (int m = 0; m < M; m++)
    if (mesh_first_triangle[m + 1] > mesh_first_triangle[m]) node_base++;
  const int nonempty = node_base;
  node_base = std::max(nonempty - 1, 0);
  for (int m = 0; m < M; m++) {
    const int t0 = mesh_first_triangle[m], n = mesh_first_triangle[m + 1] - t0;
    if (n <= 0) continue;
    roots.push_back(Seg{t0, t0 + n, node_base, -1, (int)roots.size(), 0, 0});
    node_base += n - 1;
    max_n = std::max(max_n, n);
  }
  if (nonempty < 1 || T < 2 || M > 256 || T >= (1 << 24)) {
    err = "scene outside the device builder's range";
    return false;
  }
  const int threads = 256, blocksT = (T + threads - 1) / threads;
  long long n2 = kSortTile;
  while (n2 < T) n2 <<= 1;

  {
    const char* e = getenv("RT_BVH_SMALL");  // "0": always the level-synchronous launches (tests cover both)
    if (max_n <= kSmallMeshMax && !(e && atoi(e) == 0))
      return build_small_meshes(d_pos, d_vidx, T, roots, max_n, n2, pad, d_nodes, d_tris, st, out, launches_out, err);
  }

  const int nb = (T + kScanBlock - 1) / kScanBlock;
  const size_t max_seg = (size_t)T + 2;
  Arena arena(st);
  {
    auto P = Arena::padded;
    const size_t t = (size_t)T;
    BV(arena.reserve(2 * P(t * 16) + 3 * P((size_t)n2 * 8) + 9 * P(t * 4) + 2 * P(3 * t * 4) + P(3 * (size_t)nb * 4) +
                     2 * P(4) + P(t) +
                     2 * P(max_seg * sizeof(Seg)) + 2 * P(6 * max_seg * 4) + P(6 * (size_t)nonempty * 4) +
                     P((size_t)nonempty * 4) + 4096));
  }
  struct {
    float4* p;
  } box_lo{arena.take<float4>(T)}, box_hi{arena.take<float4>(T)};
  struct {
    unsigned long long* p;
  } key[3] = {{arena.take<unsigned long long>(n2)}, {arena.take<unsigned long long>(n2)}, {arena.take<unsigned long long>(n2)}};
  struct IntBuf {
    int* p;
  };
  IntBuf ord[3] = {{arena.take<int>(T)}, {arena.take<int>(T)}, {arena.take<int>(T)}},
         ord_tmp[3] = {{arena.take<int>(T)}, {arena.take<int>(T)}, {arena.take<int>(T)}}, seg_pos{arena.take<int>(T)},
         seg_pos_next{arena.take<int>(T)}, flags{arena.take<int>(3 * (size_t)T)}, scan{arena.take<int>(3 * (size_t)T)},
         block_sum{arena.take<int>(3 * (size_t)nb)}, slot_tri{arena.take<int>(T)}, next_count{arena.take<int>(1)},
         root_ref{arena.take<int>(nonempty)};
  struct {
    unsigned char* p;
  } side{arena.take<unsigned char>(T)};
  struct {
    Seg* p;
  } seg_a{arena.take<Seg>(max_seg)}, seg_b{arena.take<Seg>(max_seg)};
  struct {
    unsigned* p;
  } box_a{arena.take<unsigned>(6 * max_seg)}, box_b{arena.take<unsigned>(6 * max_seg)};
  struct {
    float* p;
  } root_box{arena.take<float>(6 * (size_t)nonempty)};
  if (arena.used > arena.cap) {
    err = "internal: scratch arena too small";
    return false;
  }
  tick("alloc");
  // 1. keys and the three orders
  k_tri_prepare<<<blocksT, threads, 0, st>>>(d_pos, d_vidx, T, box_lo.p, box_hi.p, key[0].p, key[1].p, key[2].p);
  launches++;
  for (int a = 0; a < 3 && n2 > T; a++) {
    k_fill_u64<<<(unsigned)((n2 - T + 255) / 256), 256, 0, st>>>(key[a].p, T, n2, ~0ull);
    launches++;
  }
  if (key[1].p != key[0].p + n2 || key[2].p != key[1].p + n2 || !sort_keys(key[0].p, n2, st, &launches, 3, n2)) {
    err = "bitonic sort launch failed";  // (the three arrays are consecutive arena pieces: one batched sort)
    return false;
  }
  for (int a = 0; a < 3; a++) {
    k_extract_index<<<blocksT, threads, 0, st>>>(key[a].p, T, ord[a].p);
    launches++;
  }

  tick("keys + 3 bitonic sorts");
  // 2. level by level
  int S = nonempty;
  BV(cudaMemcpyAsync(seg_a.p, roots.data(), sizeof(Seg) * roots.size(), cudaMemcpyHostToDevice, st));
  {
    std::vector<unsigned> init(6 * (size_t)S);
    for (int s = 0; s < S; s++)
      for (int c = 0; c < 6; c++) init[6 * (size_t)s + c] = c < 3 ? 0xffffffffu : 0u;
    BV(cudaMemcpyAsync(box_a.p, init.data(), sizeof(unsigned) * init.size(), cudaMemcpyHostToDevice, st));
    BV(cudaStreamSynchronize(st));  // `init` and `roots` are pageable host memory
  }
  BV(cudaMemsetAsync(seg_pos.p, 0xff, sizeof(int) * (size_t)T, st));
  k_init_pos<<<dim3(64, S), 256, 0, st>>>(seg_a.p, S, seg_pos.p);
  launches++;
  Seg *cur = seg_a.p, *nxt = seg_b.p;
  unsigned *cur_box = box_a.p, *nxt_box = box_b.p;
  int *pos_cur = seg_pos.p, *pos_nxt = seg_pos_next.p;
  int* o[3] = {ord[0].p, ord[1].p, ord[2].p};
  int* o_tmp[3] = {ord_tmp[0].p, ord_tmp[1].p, ord_tmp[2].p};
  const int levels = tree_depth(max_n);
  // Segment counts per level are known without asking the device: a mesh of n triangles contributes segments of at
  // most two distinct sizes per level (floor and ceiling halves), so the host tracks {size: count} and the level loop
  // needs no synchronisation (15-25 levels x a stream sync + a pageable D2H was 0.7 ms of a 1.6 ms build at 11 k
  // triangles).
  std::vector<std::pair<int, long long>> sizes;  // (segment size, how many) of the current level
  for (const Seg& r : roots) sizes.push_back({r.e - r.b, 1});
  for (int level = 0; level < levels && S > 0; level++) {
    BV(cudaMemsetAsync(next_count.p, 0, sizeof(int), st));
    k_seg_box<<<blocksT, threads, 0, st>>>(pos_cur, o[0], box_lo.p, box_hi.p, T, cur_box);
    k_seg_split<<<(S + 127) / 128, 128, 0, st>>>(cur, S, cur_box, o[0], pad, (float*)d_nodes, slot_tri.p, root_box.p,
                                                root_ref.p, nxt, next_count.p, nxt_box);
    k_mark<<<blocksT, threads, 0, st>>>(cur, pos_cur, o[0], o[1], o[2], T, side.p, pos_nxt);
    launches += 3;
    if (level + 1 < levels) {  // the three orders in one launch per step (blockIdx.y = axis)
      Orders oo;
      for (int a = 0; a < 3; a++) {
        oo.in[a] = o[a];
        oo.out[a] = o_tmp[a];
      }
      k_flags<<<dim3(blocksT, 3), threads, 0, st>>>(cur, pos_cur, oo, side.p, T, flags.p);
      k_scan_block<<<dim3(nb, 3), kScanBlock, 0, st>>>(flags.p, T, scan.p, block_sum.p);
      k_scan_sums<<<3, kScanBlock, 0, st>>>(block_sum.p, nb);
      k_partition<<<dim3(blocksT, 3), threads, 0, st>>>(cur, pos_cur, oo, flags.p, scan.p, block_sum.p, T, nb);
      launches += 4;
      for (int a = 0; a < 3; a++) std::swap(o[a], o_tmp[a]);
    }
    long long S_next = 0;
    std::vector<std::pair<int, long long>> next_sizes;
    for (const auto& sc : sizes) {
      if (sc.first < 2) continue;  // leaves end here
      const int nl = sc.first / 2, nr = sc.first - nl;
      next_sizes.push_back({nl, sc.second});
      next_sizes.push_back({nr, sc.second});
      S_next += 2 * sc.second;
    }
    // merge equal sizes so the list stays at <= 2 entries per mesh
    std::sort(next_sizes.begin(), next_sizes.end());
    sizes.clear();
    for (const auto& sc : next_sizes) {
      if (!sizes.empty() && sizes.back().first == sc.first)
        sizes.back().second += sc.second;
      else
        sizes.push_back(sc);
    }
    S = (int)S_next;
    std::swap(cur, nxt);
    std::swap(cur_box, nxt_box);
    std::swap(pos_cur, pos_nxt);
  }
  if (S != 0) {
    err = "device BVH build did not terminate in the predicted number of levels";
    return false;
  }

  tick("levels");
  // 3. leaves, and the top-level join on the host
  k_leaf_tris<<<blocksT, threads, 0, st>>>(d_pos, d_vidx, slot_tri.p, T, d_tris);
  launches++;
  if (!finish_build(T, nonempty, max_n, pad, root_box.p, root_ref.p, slot_tri.p, d_nodes, st, out, err)) return false;
  if (launches_out) *launches_out = launches;
  if (timing) {
    tick("leaves + top-level join");
    fprintf(stderr, "\n");
  }
  return true;
}

}  // namespace rtb
