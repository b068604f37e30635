// kernels.cu -- the sm_100a kernels of the render hot path (wavefront formulation).
//
// Per batch of paths (nsamp samples x npix pixels) and per path segment s = 0..2:
//
//   k_raygen            (s = 0) jitter + camera ray                     Renderer.cpp:229-234
//   k_trace<NEAREST>    persistent nearest-hit BVH traversal            RayTracer.h:27-53, Ray.cpp:9-24
//   k_shade<MODE,PHOT>  hit point/normal, 3 area-light samples -> 3 shadow rays + their BSDF*radiance,
//                       hemisphere bounce -> next ray queue (compacted); or the k-nearest-photon gather
//                                                                        Renderer.cpp:33-104,143-170
//   k_trace<ANY>        persistent any-hit traversal of the shadow rays Renderer.cpp:52-55
//   k_combine           ordered sum of the unoccluded light terms, per-sample clamp on the last segment
//                                                                        Renderer.cpp:59,168,254
//   k_resolve/k_scatter ordered accumulation over samples               Renderer.cpp:255-258
//   k_emit              photon emission + Russian-roulette random walk  PhotonMap.h:14-50,92-155
//
// Why a wavefront: the first version fused a whole segment (1 nearest + 3 sequential any-hit traversals
// + shading) into one thread; ncu showed 70 % issue-slot use but only 8.1 / 6.5 of 32 lanes active on
// the bounce segments (profiles/r1_v1_k_segment_full.csv).  Here every ray is its own work item, a warp
// that runs low on live rays refills its idle lanes from the queue, and the shadow rays of 32
// neighbouring hit points towards one light share a warp.
//
// No tensor-core work exists on this path (nothing is a dense contraction): the kernels are
// pointer-chasing traversals bounded by L1/L2 latency and the fp32/ALU issue rate.
#include <limits.h>
#include <stdlib.h>

#include "kernels.h"

namespace rtb {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kDone = INT_MIN;    // traversal cursor value: no work (never a valid ~slot)
#ifndef RT_TRACE_MINB
#define RT_TRACE_MINB 10  // minimum CTAs/SM of the postponed-leaf trace kernels: 47 registers, no spills (1: 55 registers, 9 CTAs; 12: 40 registers, slower)
#endif
#ifndef RT_SHADE_MINB
#define RT_SHADE_MINB 8  // minimum CTAs/SM of the direct-lighting k_shade (register cap 64)
#endif
#ifndef RT_ANY_FIXED_ORDER
#define RT_ANY_FIXED_ORDER 0
#endif
constexpr bool kAnyFixedOrder = RT_ANY_FIXED_ORDER != 0;  // any-hit: visit child 0 first instead of the nearer child
constexpr int kRefillBelowDefault = 14;  // refill a warp's idle lanes when fewer lanes than this are live
__constant__ int c_refill_below = kRefillBelowDefault;          // any-hit launches
__constant__ int c_refill_below_nearest = kRefillBelowDefault;  // nearest-hit launches (RT_REFILL_BELOW_NEAREST)
#define kRefillBelow (ANY ? c_refill_below : c_refill_below_nearest)

// ----------------------------------------------------------------------------------------------
// k_trace: persistent warps, one ray per lane, lanes refilled from the queue as their rays finish.
// Shared memory: per thread a stack of (node ref, tnear) pairs, [depth][kBlock] ints each.
//
// Tuning knobs (rt_set_tuning / RT_TUNING env, measured on B200 -- see DESIGN.md section 4):
//   LEAFB  > 0: lanes that reach a leaf wait until at least LEAFB lanes of the warp hold a leaf (or no
//               lane has an internal node left), so the Moller-Trumbore block is issued for many lanes
//               at once instead of once per iteration for a few ("while-while" with a vote);
//   MINB      : __launch_bounds__ minimum CTAs per SM (register cap -> occupancy).
// ----------------------------------------------------------------------------------------------
struct TraceLane {
  unsigned idx;
  float3 o, d, inv, oinv;
  HitRec h;
  int sp, cur;
};

// Entry distance of a stacked subtree: only used to cull it at pop time (`tnear <= best t`), so it is kept as the upper
// 16 bits of the binary32 (tnear >= 0: truncation rounds DOWN, the test stays conservative and the hits exact).  Six
// bytes per stack entry instead of eight: 30 -> 22.5 KB per CTA on the 1.23 M-triangle scene (7 -> 9 CTAs per SM).
typedef unsigned short TnT;
RT_DI TnT tn_store(float t) { return (TnT)(__float_as_uint(t) >> 16); }
RT_DI float tn_load(TnT v) { return __uint_as_float((unsigned)v << 16); }

template <bool ANY>
RT_DI void trace_pop(TraceLane& L, const int* st_ref, const TnT* st_tn) {
  L.cur = kDone;
  while (L.sp > 0) {  // next entry whose box can still matter
    L.sp--;
    const int ref = st_ref[L.sp * kBlock];
    if (ANY || tn_load(st_tn[L.sp * kBlock]) <= L.h.t) {
      L.cur = ref;
      break;
    }
  }
}
template <bool ANY>
RT_DI void trace_write(const TraceLane& L, float4* __restrict__ hits, unsigned char* __restrict__ occ) {
  if (ANY)
    __stcs(occ + L.idx, (unsigned char)((L.h.gid != 0x7fffffff) ? 1 : 0));
  else
    __stcs(hits + L.idx, make_float4(L.h.t, L.h.u, L.h.v, __int_as_float(L.h.gid != 0x7fffffff ? L.h.gid : -1)));
}
// A new ray enters the scene through the list of per-mesh roots (the reference loops over the meshes,
// RayTracer.h:56-85): all root boxes are tested back to back, the nearest hit one becomes the cursor and the others
// go on the stack with their entry distance.  ncu on the version that entered through the top-level tree showed
// 12.1 node visits per shadow ray, about 6 of them in the top-level tree whose wall-sized boxes almost never cull.
template <bool ANY>
RT_DI void trace_start(const DScene& S, TraceLane& L, int* st_ref, TnT* st_tn) {
  L.sp = 0;
  if (S.num_roots == 0) {
    L.cur = 0;
    return;
  }
  L.cur = kDone;
  float tcur = 0.f;
  for (int r = 0; r < S.num_roots; r++) {
    const float4 lo = S.root_lo[r], hi = S.root_hi[r];
    float tn;
    if (box_hit_fma<true>(f3(lo), f3(hi), L.inv, L.oinv, FLT_MAX, tn)) {
      const int ref = __float_as_int(hi.w);
      if (L.cur == kDone) {
        L.cur = ref;
        tcur = tn;
      } else {
        const bool nearer = !ANY && tn < tcur;
        st_ref[L.sp * kBlock] = nearer ? L.cur : ref;
        if (!ANY) st_tn[L.sp * kBlock] = tn_store(nearer ? tcur : tn);
        L.sp++;
        if (nearer) {
          L.cur = ref;
          tcur = tn;
        }
      }
    }
  }
}

// internal node: test both children, descend into the nearer hit one, push the other
template <bool ANY>
RT_DI void trace_box_step(const DScene& S, TraceLane& L, int* st_ref, TnT* st_tn) {
  const float4 n0 = __ldg(S.nodes + 4 * L.cur), n1 = __ldg(S.nodes + 4 * L.cur + 1);
  const float4 n2 = __ldg(S.nodes + 4 * L.cur + 2), n3 = __ldg(S.nodes + 4 * L.cur + 3);
  float tn0, tn1;
  const bool h0 = box_hit_fma<ANY>(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), L.inv, L.oinv, L.h.t, tn0);
  const bool h1 = box_hit_fma<ANY>(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), L.inv, L.oinv, L.h.t, tn1);
  const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
  if (h0 && h1) {
    const bool first0 = tn0 <= tn1;
    st_ref[L.sp * kBlock] = first0 ? c1 : c0;
    if (!ANY) st_tn[L.sp * kBlock] = tn_store(first0 ? tn1 : tn0);
    L.sp++;
    L.cur = first0 ? c0 : c1;
  } else if (h0) {
    L.cur = c0;
  } else if (h1) {
    L.cur = c1;
  } else {
    trace_pop<ANY>(L, st_ref, st_tn);
  }
}
// leaf: one triangle.  Returns true when an any-hit ray is finished by this triangle.
template <bool ANY>
RT_DI void trace_leaf_step(const DScene& S, TraceLane& L, const int* st_ref, const TnT* st_tn) {
  const int slot = ~L.cur;
  const float4 A = __ldg(S.tris + 3 * slot), B = __ldg(S.tris + 3 * slot + 1), C = __ldg(S.tris + 3 * slot + 2);
  float u, v, t;
  bool stop = false;
  if (mt_intersect(L.o, L.d, f3(A), f3(B), f3(C), u, v, t)) {
    if (ANY) {
      if (t > 0.f && t < FLT_MAX) {
        stop = true;  // occluded: this ray is done
        L.h.gid = 0;
      }
    } else {
      accept_nearest(L.h, t, u, v, __float_as_int(A.w));
    }
  }
  if (stop)
    L.cur = kDone;
  else
    trace_pop<ANY>(L, st_ref, st_tn);
}

// Next work item of a trace kernel.  Plain queues hold (origin, direction) per ray.  BLOCKED (shadow rays): slot `my`
// of the light-major layout belongs to hit j = shadow_slot_hit(my, nl); its origin is the hit point hit_p[j] shared by
// the hit's nl rays, its direction rd[my]; w != 0 marks a ray k_shade already found occluded by the very triangle it
// starts on (12-15 % of all shadow rays) -- the result is written, nothing to trace.
template <bool BLOCKED>
RT_DI bool trace_fetch(const float4* __restrict__ ro, const float4* __restrict__ rd, unsigned my, unsigned n,
                       unsigned items, unsigned nl, float4& a, float4& b) {
  if (my >= n) return false;
  if (BLOCKED) {
    const unsigned j = shadow_slot_hit(my, nl);
    if (j >= items) return false;
    b = __ldcs(rd + my);  // streamed once: keep L1 for the BVH
    if (b.w != 0.f) return false;
    a = __ldg(ro + j);
    return true;
  }
  a = __ldcs(ro + my);
  b = __ldcs(rd + my);
  return true;
}

template <bool ANY, bool BLOCKED, int LEAFB>
RT_DI void trace_body(const DScene& S, const float4* __restrict__ ro, const float4* __restrict__ rd,
                      const unsigned* n_ptr, unsigned n_fixed, float4* __restrict__ hits,
                      unsigned char* __restrict__ occ, unsigned* fetch, int depth, unsigned nl, int* s_dyn) {
  int* st_ref = s_dyn + threadIdx.x;
  TnT* st_tn = reinterpret_cast<TnT*>(s_dyn + depth * kBlock) + threadIdx.x;
  const unsigned items = n_ptr ? *n_ptr : n_fixed;                  // rays (or hit points when BLOCKED)
  const unsigned n = BLOCKED ? (unsigned)shadow_slots_for(items, nl) : items;  // queue slots to visit
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;

  TraceLane L;
  L.idx = 0;
  L.o = L.d = L.inv = L.oinv = f3(0, 0, 0);
  L.h.t = FLT_MAX;
  L.h.u = L.h.v = 0.f;
  L.h.gid = 0x7fffffff;
  L.sp = 0;
  L.cur = kDone;
  bool exhausted = false;  // warp-uniform

  for (;;) {
    // ---- refill idle lanes -------------------------------------------------------------------
    const bool need = (L.cur == kDone);
    const unsigned need_mask = __ballot_sync(kFull, need);
    if (!exhausted && need_mask) {
      const unsigned cnt = __popc(need_mask);
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(fetch, cnt);
      base = __shfl_sync(kFull, base, 0);
      if (base + cnt >= n) exhausted = true;
      if (need) {
        const unsigned my = base + __popc(need_mask & lt_mask);
        float4 a, b;
        if (trace_fetch<BLOCKED>(ro, rd, my, n, items, nl, a, b)) {
          L.idx = my;
          L.o = f3(a);
          L.d = f3(b);
          L.inv = f3(safe_inv(L.d.x), safe_inv(L.d.y), safe_inv(L.d.z));
          L.oinv = f3(L.o.x * L.inv.x, L.o.y * L.inv.y, L.o.z * L.inv.z);
          L.h.t = FLT_MAX;
          L.h.u = L.h.v = 0.f;
          L.h.gid = 0x7fffffff;
          trace_start<ANY>(S, L, st_ref, st_tn);
          if (L.cur == kDone) trace_write<ANY>(L, hits, occ);  // no mesh box hit: a miss
        }
      }
    }
    if (__ballot_sync(kFull, L.cur != kDone) == 0) {
      if (exhausted) break;
      continue;
    }
    // ---- traverse until the warp runs low on live rays ------------------------------------------
    for (;;) {
      if (LEAFB == 0) {
        if (L.cur != kDone) {
          if (L.cur >= 0)
            trace_box_step<ANY>(S, L, st_ref, st_tn);
          else
            trace_leaf_step<ANY>(S, L, st_ref, st_tn);
          if (L.cur == kDone) trace_write<ANY>(L, hits, occ);  // ray complete, the lane becomes idle
        }
      } else {
        if (L.cur >= 0) {
          trace_box_step<ANY>(S, L, st_ref, st_tn);
          if (L.cur == kDone) trace_write<ANY>(L, hits, occ);
        }
        const bool at_leaf = L.cur < 0 && L.cur != kDone;
        const unsigned leaf_mask = __ballot_sync(kFull, at_leaf);
        if (leaf_mask) {
          const unsigned node_mask = __ballot_sync(kFull, L.cur >= 0);
          if (__popc(leaf_mask) >= LEAFB || node_mask == 0) {
            if (at_leaf) {
              trace_leaf_step<ANY>(S, L, st_ref, st_tn);
              if (L.cur == kDone) trace_write<ANY>(L, hits, occ);
            }
          }
        }
      }
      const unsigned live = __ballot_sync(kFull, L.cur != kDone);
      if (live == 0) break;
      if (!exhausted && __popc(live) < kRefillBelow) break;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// trace_body_spec: the same persistent loop with POSTPONED leaves.  ncu on the batched-leaf body
// (profiles/r1_v2b_leafbatch_full.csv, SASS view) showed the box step running with 24 of 32 lanes (the others
// sit on a leaf waiting for the batch), the Moller-Trumbore block with 12, and two separate copies of the stack
// pop loop with 3 and 8 lanes.  Here a lane that reaches a leaf parks it in `pend` and keeps traversing; the
// triangle block runs when >= LEAFT lanes have a parked leaf (or no lane has a node left), and there is ONE pop
// site shared by "both children missed" and "leaf parked".  The result does not depend on the order in which
// leaves are tested (accept_nearest is order-independent; any-hit is an OR), only the culling is a little later.
// ----------------------------------------------------------------------------------------------
template <bool ANY, bool BLOCKED, int LEAFT, int UNROLL = 1>
RT_DI void trace_body_spec(const DScene& S, const float4* __restrict__ ro, const float4* __restrict__ rd,
                           const unsigned* n_ptr, unsigned n_fixed, float4* __restrict__ hits,
                           unsigned char* __restrict__ occ, unsigned* fetch, int depth, unsigned nl, int* s_dyn) {
  int* st_ref = s_dyn + threadIdx.x;
  TnT* st_tn = reinterpret_cast<TnT*>(s_dyn + depth * kBlock) + threadIdx.x;
  const unsigned items = n_ptr ? *n_ptr : n_fixed;
  const unsigned n = BLOCKED ? (unsigned)shadow_slots_for(items, nl) : items;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;

  TraceLane L;
  L.idx = 0;
  L.o = L.d = L.inv = L.oinv = f3(0, 0, 0);
  L.h.t = FLT_MAX;
  L.h.u = L.h.v = 0.f;
  L.h.gid = 0x7fffffff;
  L.sp = 0;
  L.cur = kDone;
  int pend = 0;       // parked leaf reference (< 0) or 0
  bool live = false;  // this lane owns an unfinished ray
  bool exhausted = false;

  for (;;) {
    const unsigned need_mask = __ballot_sync(kFull, !live);
    if (!exhausted && need_mask) {
      const unsigned cnt = __popc(need_mask);
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(fetch, cnt);
      base = __shfl_sync(kFull, base, 0);
      if (base + cnt >= n) exhausted = true;
      if (!live) {
        const unsigned my = base + __popc(need_mask & lt_mask);
        float4 a, b;
        if (trace_fetch<BLOCKED>(ro, rd, my, n, items, nl, a, b)) {
          L.idx = my;
          L.o = f3(a);
          L.d = f3(b);
          L.inv = f3(safe_inv(L.d.x), safe_inv(L.d.y), safe_inv(L.d.z));
          L.oinv = f3(L.o.x * L.inv.x, L.o.y * L.inv.y, L.o.z * L.inv.z);
          L.h.t = FLT_MAX;
          L.h.u = L.h.v = 0.f;
          L.h.gid = 0x7fffffff;
          trace_start<ANY>(S, L, st_ref, st_tn);
          pend = 0;
          live = true;
        }
      }
    }
    if (__ballot_sync(kFull, live) == 0) {
      if (exhausted) break;
      continue;
    }
    for (;;) {
#pragma unroll
     for (int u = 0; u < UNROLL; u++) {  // UNROLL node steps per round of votes (the votes cost ~26 issue slots)
      bool need_pop = false;
      if (L.cur >= 0) {  // internal node: both child boxes
        const float4 n0 = __ldg(S.nodes + 4 * L.cur), n1 = __ldg(S.nodes + 4 * L.cur + 1);
        const float4 n2 = __ldg(S.nodes + 4 * L.cur + 2), n3 = __ldg(S.nodes + 4 * L.cur + 3);
        float tn0, tn1;
        const bool h0 = box_hit_fma<ANY>(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), L.inv, L.oinv, L.h.t, tn0);
        const bool h1 = box_hit_fma<ANY>(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), L.inv, L.oinv, L.h.t, tn1);
        const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
        const bool first0 = (ANY && kAnyFixedOrder) ? true : tn0 <= tn1;
        if (h0 && h1) {
          st_ref[L.sp * kBlock] = first0 ? c1 : c0;
          if (!ANY) st_tn[L.sp * kBlock] = tn_store(first0 ? tn1 : tn0);
          L.sp++;
          L.cur = first0 ? c0 : c1;
        } else if (h0 || h1) {
          L.cur = h0 ? c0 : c1;
        } else {
          need_pop = true;
        }
      }
      if (!need_pop && L.cur < 0 && L.cur != kDone && pend == 0) {  // park the leaf, keep traversing
        pend = L.cur;
        need_pop = true;
      }
      if (need_pop) trace_pop<ANY>(L, st_ref, st_tn);
     }

      const unsigned pend_mask = __ballot_sync(kFull, pend != 0);
      if (pend_mask) {
        const unsigned node_mask = __ballot_sync(kFull, L.cur >= 0);
        if (__popc(pend_mask) >= LEAFT || node_mask == 0) {
          if (pend != 0) {
            const int slot = ~pend;
            pend = 0;
            const float4 A = __ldg(S.tris + 3 * slot), B = __ldg(S.tris + 3 * slot + 1), C = __ldg(S.tris + 3 * slot + 2);
            float u, v, t;
            if (mt_intersect(L.o, L.d, f3(A), f3(B), f3(C), u, v, t)) {
              if (ANY) {
                if (t > 0.f && t < FLT_MAX) {
                  L.h.gid = 0;
                  L.cur = kDone;  // occluded: drop the rest of the traversal
                  L.sp = 0;
                }
              } else {
                accept_nearest(L.h, t, u, v, __float_as_int(A.w));
              }
            }
          }
        }
      }
      if (live && L.cur == kDone && pend == 0) {
        trace_write<ANY>(L, hits, occ);
        live = false;
      }
      const unsigned live_mask = __ballot_sync(kFull, live);
      if (live_mask == 0) break;
      if (!exhausted && __popc(live_mask) < kRefillBelow) break;
    }
  }
}

#define RT_TRACE_KERNEL_SPEC(NAME, LEAFT, UNROLL)                                                                           \
  template <bool ANY, bool BLOCKED>                                                                                 \
  __global__ void __launch_bounds__(kBlock, RT_TRACE_MINB)                                                                      \
      NAME(const DScene S, const float4* __restrict__ ro, const float4* __restrict__ rd, const unsigned* n_ptr,    \
           unsigned n_fixed, float4* __restrict__ hits, unsigned char* __restrict__ occ, unsigned* fetch, int depth, \
           unsigned nl) {                                                                                           \
    extern __shared__ int s_dyn[];                                                                                  \
    trace_body_spec<ANY, BLOCKED, LEAFT, UNROLL>(S, ro, rd, n_ptr, n_fixed, hits, occ, fetch, depth, nl, s_dyn);    \
  }
// Variants kept for comparison (RT_TRACE_VARIANT); everything else that was measured is in profiles/r1_tuning.md.
RT_TRACE_KERNEL_SPEC(k_trace_sp8, 8, 1)    // variant 5: one node step per round of votes (28.1 ms/frame)
RT_TRACE_KERNEL_SPEC(k_trace_sp8u3, 8, 3)  // variant 10 (default): 26.2 ms/frame; u2 26.7, u4 26.8, u6 27.7, u1 28.1

#define RT_TRACE_KERNEL(NAME, LEAFB, MINB)                                                                        \
  template <bool ANY, bool BLOCKED>                                                                                 \
  __global__ void __launch_bounds__(kBlock, MINB)                                                                   \
      NAME(const DScene S, const float4* __restrict__ ro, const float4* __restrict__ rd, const unsigned* n_ptr,    \
           unsigned n_fixed, float4* __restrict__ hits, unsigned char* __restrict__ occ, unsigned* fetch, int depth, \
           unsigned nl) {                                                                                           \
    extern __shared__ int s_dyn[];                                                                                  \
    trace_body<ANY, BLOCKED, LEAFB>(S, ro, rd, n_ptr, n_fixed, hits, occ, fetch, depth, nl, s_dyn);                 \
  }
// Measured on B200, cfg2 frame (profiles/r1_tuning.md): variant 0 37.4 ms, variant 1 33.6 ms at the time; the
// postponed-leaf kernels above superseded both.
RT_TRACE_KERNEL(k_trace_lb8, 8, 1)    // variant 1: batch leaf tests, >= 8 lanes
RT_TRACE_KERNEL(k_trace, 0, 1)        // variant 0: test leaves as they come

static int g_trace_variant = -1;
int trace_variant() {
  if (g_trace_variant < 0) {
    const char* e = getenv("RT_TRACE_VARIANT");
    g_trace_variant = e ? atoi(e) : 10;
  }
  return g_trace_variant;
}
void set_trace_variant(int v) { g_trace_variant = v; }

// RayTracer.h:27-53 exactly as written: every ray scans every triangle (parity hook, RT_FLAG_BRUTE_FORCE)
template <bool ANY, bool BLOCKED>
__global__ void __launch_bounds__(kBlock)
    k_trace_brute(const DScene S, const float4* __restrict__ ro, const float4* __restrict__ rd, const unsigned* n_ptr,
                  unsigned n_fixed, float4* __restrict__ hits, unsigned char* __restrict__ occ, unsigned nl) {
  const unsigned items = n_ptr ? *n_ptr : n_fixed;
  const unsigned n = BLOCKED ? (unsigned)shadow_slots_for(items, nl) : items;
  for (unsigned i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    float4 a, b;
    if (!trace_fetch<BLOCKED>(ro, rd, i, n, items, nl, a, b)) continue;
    HitRec h;
    const float3 o = f3(a), d = f3(b);
    const bool f = brute_trace<ANY>(S, o, d, h);
    if (ANY)
      occ[i] = f ? 1 : 0;
    else
      hits[i] = make_float4(h.t, h.u, h.v, __int_as_float(f ? h.gid : -1));
  }
}

// per stack entry and thread: a 4-byte reference, and for nearest-hit a 2-byte entry distance (TnT)
size_t trace_smem_bytes(int stack_depth, bool any) {
  return (size_t)stack_depth * kBlock * (sizeof(int) + (any ? 0 : sizeof(TnT)));
}
size_t trace_smem_bytes(int stack_depth) { return trace_smem_bytes(stack_depth, false); }

template <bool ANY>
static int trace_occupancy(int stack_depth) {
  int n = 0;
  const size_t sm = trace_smem_bytes(stack_depth, ANY);
  switch (trace_variant()) {
    case 0: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace<ANY, ANY>, kBlock, sm); break;
    case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace_lb8<ANY, ANY>, kBlock, sm); break;
    case 5: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace_sp8<ANY, ANY>, kBlock, sm); break;
    default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_trace_sp8u3<ANY, ANY>, kBlock, sm); break;
  }
  return n < 1 ? 1 : n;
}
// resident CTAs per SM of the nearest-hit kernel / of the any-hit kernel (no entry-distance column: smaller stack)
int trace_ctas_per_sm(int stack_depth) { return trace_occupancy<false>(stack_depth); }
int trace_any_ctas_per_sm(int stack_depth) { return trace_occupancy<true>(stack_depth); }

template <bool ANY, bool BLOCKED>
static void launch_trace_t(const DScene& S, const float4* ro, const float4* rd, const unsigned* n_ptr, unsigned n_fixed,
                           float4* hits, unsigned char* occ, unsigned* fetch, int brute, int depth, unsigned nl,
                           int grid, cudaStream_t st) {
  if (brute) {
    k_trace_brute<ANY, BLOCKED><<<grid, kBlock, 0, st>>>(S, ro, rd, n_ptr, n_fixed, hits, occ, nl);
  } else {
    cudaMemsetAsync(fetch, 0, sizeof(unsigned), st);
    static const int refill = getenv("RT_REFILL_BELOW") ? atoi(getenv("RT_REFILL_BELOW")) : -1;
    static bool refill_set = false;
    static const int refill_nearest =
        getenv("RT_REFILL_BELOW_NEAREST") ? atoi(getenv("RT_REFILL_BELOW_NEAREST")) : refill;
    if ((refill >= 0 || refill_nearest >= 0) && !refill_set) {
      if (refill >= 0) cudaMemcpyToSymbol(c_refill_below, &refill, sizeof(int));
      if (refill_nearest >= 0) cudaMemcpyToSymbol(c_refill_below_nearest, &refill_nearest, sizeof(int));
      refill_set = true;
    }
    // any-hit keeps no tnear column: half the stack, and the rest of the SM's 256 KB stays L1 for the BVH
    const size_t sm = trace_smem_bytes(depth, ANY);
    static const int carve = getenv("RT_CARVEOUT") ? atoi(getenv("RT_CARVEOUT")) : -1;
#define RT_LAUNCH(K)                                                                                          \
  do {                                                                                                        \
    if (carve >= 0) cudaFuncSetAttribute(K<ANY, BLOCKED>, cudaFuncAttributePreferredSharedMemoryCarveout, carve); \
    K<ANY, BLOCKED><<<grid, kBlock, sm, st>>>(S, ro, rd, n_ptr, n_fixed, hits, occ, fetch, depth, nl);       \
  } while (0)
    switch (trace_variant()) {
      case 0: RT_LAUNCH(k_trace); break;
      case 1: RT_LAUNCH(k_trace_lb8); break;
      case 5: RT_LAUNCH(k_trace_sp8); break;
      default: RT_LAUNCH(k_trace_sp8u3); break;
    }
#undef RT_LAUNCH
  }
}

void launch_trace_nearest(const RenderArgs& a, int seg, int grid, cudaStream_t st) {
  const unsigned* n_ptr = seg == 0 ? nullptr : a.q_count + kQHits0 + (seg - 1);
  launch_trace_t<false, false>(a.scene, a.ray_o[seg & 1], a.ray_d[seg & 1], n_ptr, (unsigned)a.npix * (unsigned)a.nsamp,
                               a.hit, nullptr, a.q_count + kQFetchNearest0 + seg, a.brute, a.stack_depth, 1u, grid, st);
}
void launch_trace_any(const RenderArgs& a, int seg, int grid, cudaStream_t st) {
  if (a.nl < 1) return;  // a scene without lights casts no shadow rays
  launch_trace_t<true, true>(a.scene, a.hit_p, a.sh_d, a.q_count + kQHits0 + seg, 0, nullptr, a.occ,
                             a.q_count + kQFetchAny0 + seg, a.brute, a.stack_depth, (unsigned)a.nl, grid, st);
}
void launch_trace_rays(const DScene& s, const float4* ro, const float4* rd, unsigned n, float4* hits,
                       unsigned char* occluded, int any, int brute, int stack_depth, unsigned* fetch_counter, int grid,
                       cudaStream_t st) {
  if (any)
    launch_trace_t<true, false>(s, ro, rd, nullptr, n, nullptr, occluded, fetch_counter, brute, stack_depth, 1u, grid, st);
  else
    launch_trace_t<false, false>(s, ro, rd, nullptr, n, hits, nullptr, fetch_counter, brute, stack_depth, 1u, grid, st);
}

// ----------------------------------------------------------------------------------------------
// k_raygen: path p -> primary ray (Renderer.cpp:229-234)
// ----------------------------------------------------------------------------------------------
RT_DI uint64_t path_stream_key(const RenderArgs& A, unsigned p, int& pixel, int& sample) {
  const int sl = p / (unsigned)A.npix;
  const int pl = p - sl * A.npix;
  pixel = __ldg(A.pix_map + pl);
  sample = A.s0 + sl;
  return stream_key(A.seed_mixed, kDomainPixel, (uint64_t)sample * ((uint64_t)A.width * A.height) + (uint64_t)pixel);
}

__global__ void __launch_bounds__(256) k_raygen(const RenderArgs A) {
  const unsigned n = (unsigned)A.npix * (unsigned)A.nsamp;
  for (unsigned p = blockIdx.x * 256 + threadIdx.x; p < n; p += gridDim.x * 256) {
    int pixel, sample;
    Rng g;
    g.init(path_stream_key(A, p, pixel, sample), 0);
    float sx, sy;
    jitter_sample(g, sample, A.jitter_d, sx, sy);
    const int y = pixel / A.width, x = pixel - y * A.width;
    float3 o, d;
    camera_ray(A.scene.cam, x, y, sx, sy, A.width, A.height, o, d);
    A.ray_o[0][p] = make_float4(o.x, o.y, o.z, __int_as_float((int)p));
    A.ray_d[0][p] = make_float4(d.x, d.y, d.z, 0.f);
  }
}
void launch_raygen(const RenderArgs& a, cudaStream_t st) {
  long long n = (long long)a.npix * a.nsamp;
  int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (grid < 1) grid = 1;
  k_raygen<<<grid, 256, 0, st>>>(a);
}

// ----------------------------------------------------------------------------------------------
// Spatial binning of a bounce segment's hit points.  After a diffuse bounce the hit points of a warp's
// 32 rays are scattered over the scene, and so are the 3 shadow rays each of them spawns (any-hit lanes
// 18.6 of 32 on segment 1 against 22.7 on the pixel-coherent segment 0, profiles/r1_v2b_leafbatch_full.csv).
// A counting sort by the Morton code of the hit point's cell restores that coherence for the price of
// three small memory-bound kernels.  The order inside a cell is whatever the atomics produce; results do
// not depend on it (every path owns its random stream and its output slot).
// ----------------------------------------------------------------------------------------------
RT_DI unsigned morton3_5bit(unsigned x, unsigned y, unsigned z) {
  auto spread = [](unsigned v) {  // 5 bits -> every third bit
    v = (v | (v << 8)) & 0x0000F00Fu;
    v = (v | (v << 4)) & 0x000C30C3u;
    v = (v | (v << 2)) & 0x00249249u;
    return v;
  };
  return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}
RT_DI unsigned sort_key(const RenderArgs& A, int seg, unsigned i) {
  const float4 hr = A.hit[i];
  if (__float_as_int(hr.w) < 0) return (unsigned)kSortBuckets;  // misses last
  const float4 o = A.ray_o[seg & 1][i], d = A.ray_d[seg & 1][i];
  const float px = fmaf(hr.x, d.x, o.x), py = fmaf(hr.x, d.y, o.y), pz = fmaf(hr.x, d.z, o.z);  // ~hit point
  const int cx = min(max((int)((px - A.sort_lo.x) * A.sort_inv_cell.x), 0), kSortGrid - 1);
  const int cy = min(max((int)((py - A.sort_lo.y) * A.sort_inv_cell.y), 0), kSortGrid - 1);
  const int cz = min(max((int)((pz - A.sort_lo.z) * A.sort_inv_cell.z), 0), kSortGrid - 1);
  return morton3_5bit((unsigned)cx, (unsigned)cy, (unsigned)cz);
}
// One CTA of 1024 threads per SM; CTA c owns the contiguous chunk [c*chunk, (c+1)*chunk) of the ray queue in
// BOTH passes.  Shared memory holds one counter per bucket (128 KB + 4 B): hot cells (a wall) and the miss
// bucket are absorbed by shared-memory atomics, and each CTA touches a global counter once per non-empty bucket.
constexpr int kSortThreads = 1024;
constexpr int kSortCells = kSortBuckets + 1;
constexpr size_t kSortSmem = (size_t)kSortCells * sizeof(unsigned);

__global__ void __launch_bounds__(kSortThreads) k_sort_count(const RenderArgs A, const int seg) {
  extern __shared__ unsigned s_cnt[];
  const unsigned n = seg == 0 ? (unsigned)A.npix * (unsigned)A.nsamp : A.q_count[kQHits0 + seg - 1];
  const unsigned chunk = (n + gridDim.x - 1) / gridDim.x;
  const unsigned i0 = blockIdx.x * chunk, i1 = min(n, i0 + chunk);
  for (int b = threadIdx.x; b < kSortCells; b += kSortThreads) s_cnt[b] = 0;
  __syncthreads();
  for (unsigned i = i0 + threadIdx.x; i < i1; i += kSortThreads) {
    const unsigned key = sort_key(A, seg, i);
    A.perm[i] = key;  // the scatter pass reads 4 bytes per ray instead of recomputing the key from 48
    atomicAdd(s_cnt + key, 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kSortCells; b += kSortThreads)
    if (s_cnt[b]) atomicAdd(A.sort_hist + b, s_cnt[b]);
}
// exclusive scan of kSortBuckets+1 counters in place (one CTA of 1024 threads, 33 counters per thread)
__global__ void __launch_bounds__(1024) k_sort_scan(unsigned* hist) {
  __shared__ unsigned s_part[1024];
  constexpr int kPer = (kSortCells + 1023) / 1024;
  unsigned local[kPer];
  unsigned sum = 0;
  const int first = threadIdx.x * kPer;
  for (int k = 0; k < kPer; k++) {
    const int b = first + k;
    local[k] = b < kSortCells ? hist[b] : 0u;
    sum += local[k];
  }
  s_part[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan of the partials
    unsigned v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0u;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned run = s_part[threadIdx.x] - sum;
  for (int k = 0; k < kPer; k++) {
    const int b = first + k;
    if (b < kSortCells) hist[b] = run;
    run += local[k];
  }
}
__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const RenderArgs A, const int seg) {
  extern __shared__ unsigned s_cnt[];
  const unsigned n = seg == 0 ? (unsigned)A.npix * (unsigned)A.nsamp : A.q_count[kQHits0 + seg - 1];
  const unsigned chunk = (n + gridDim.x - 1) / gridDim.x;
  const unsigned i0 = blockIdx.x * chunk, i1 = min(n, i0 + chunk);
  for (int b = threadIdx.x; b < kSortCells; b += kSortThreads) s_cnt[b] = 0;
  __syncthreads();
  for (unsigned i = i0 + threadIdx.x; i < i1; i += kSortThreads) atomicAdd(s_cnt + A.perm[i], 1u);
  __syncthreads();
  for (int b = threadIdx.x; b < kSortCells; b += kSortThreads)  // reserve this CTA's range in every bucket
    if (s_cnt[b]) s_cnt[b] = atomicAdd(A.sort_hist + b, s_cnt[b]);
  __syncthreads();
  // The payload k_shade needs (hit record, ray direction, path id) moves to its sorted slot here, where it has
  // just been read in queue order for the key: random 16-byte WRITES that nobody waits for, instead of random
  // reads in k_shade (ncu, profiles/r1c_cfg2_frame_full.csv: 7.3 GB of DRAM reads and 47 % issue activity in the
  // gathering k_shade of a bounce segment against 1.1 GB / 68 % on the pixel-ordered segment 0).
  for (unsigned i = i0 + threadIdx.x; i < i1; i += kSortThreads) {
    const unsigned dst = atomicAdd(s_cnt + A.perm[i], 1u);
    const float4 o = A.ray_o[seg & 1][i], d = A.ray_d[seg & 1][i];
    A.sorted[2 * (size_t)dst] = A.hit[i];  // one full 32-byte sector per ray
    A.sorted[2 * (size_t)dst + 1] = make_float4(d.x, d.y, d.z, o.w);
  }
}
// ---- the same binning with 2^bits cells per axis (bits = 6, 7: 262 144 / 2 097 152 Morton cells) ----------------------
// Finer cells put rays that start almost at the same point into the same warp (their traversals -- and their k-NN
// queries -- then run in lock-step), but the counters no longer fit in shared memory: the histogram lives in global
// memory (L2-resident, <= 8 MB), both passes are plain grid-stride kernels, and only the miss bucket -- the one address
// every lane of a warp can want -- is warp-aggregated.  The scan is two kernels (per-block exclusive scan + block
// sums, scan of the block sums); the scatter adds the block offset itself.
RT_DI unsigned morton3_10bit(unsigned x, unsigned y, unsigned z) {
  auto spread = [](unsigned v) {  // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
  };
  return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}
RT_DI unsigned sort_key_fine(const RenderArgs& A, int seg, unsigned i) {
  const float4 hr = A.hit[i];
  if (__float_as_int(hr.w) < 0) return 1u << (3 * A.sort_bits);  // misses last
  const float4 o = A.ray_o[seg & 1][i], d = A.ray_d[seg & 1][i];
  const float px = fmaf(hr.x, d.x, o.x), py = fmaf(hr.x, d.y, o.y), pz = fmaf(hr.x, d.z, o.z);  // ~hit point
  const int g1 = (1 << A.sort_bits) - 1;
  const int cx = min(max((int)((px - A.sort_lo.x) * A.sort_inv_cell.x), 0), g1);
  const int cy = min(max((int)((py - A.sort_lo.y) * A.sort_inv_cell.y), 0), g1);
  const int cz = min(max((int)((pz - A.sort_lo.z) * A.sort_inv_cell.z), 0), g1);
  return morton3_10bit((unsigned)cx, (unsigned)cy, (unsigned)cz);
}
constexpr int kScanThreads = 1024, kScanPer = 8, kScanTile = kScanThreads * kScanPer;
__global__ void k_sort_count_fine(const RenderArgs A, const int seg) {
  const unsigned n = seg == 0 ? (unsigned)A.npix * (unsigned)A.nsamp : A.q_count[kQHits0 + seg - 1];
  const unsigned miss = 1u << (3 * A.sort_bits);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned key = sort_key_fine(A, seg, i);
    A.perm[i] = key;  // the scatter pass reads 4 bytes per ray instead of recomputing the key from 48
    if (key == miss) {
      const unsigned m = __activemask();
      if ((threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(A.sort_hist + miss, (unsigned)__popc(m));
    } else {
      atomicAdd(A.sort_hist + key, 1u);
    }
  }
}
// in place: hist[i] <- exclusive prefix inside its tile of kScanTile counters; tile_sum[tile] <- the tile's total
__global__ void __launch_bounds__(kScanThreads) k_sort_scan_tiles(unsigned* hist, unsigned n, unsigned* tile_sum) {
  __shared__ unsigned s_warp[kScanThreads / 32];
  const unsigned first = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
  unsigned v[kScanPer], sum = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; k++) {
    v[k] = first + k < n ? hist[first + k] : 0u;
    sum += v[k];
  }
  unsigned incl = sum;
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
    if ((threadIdx.x & 31) >= off) incl += o;
  }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    const unsigned w = s_warp[threadIdx.x];
    unsigned wi = w;
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, wi, off);
      if (threadIdx.x >= off) wi += o;
    }
    s_warp[threadIdx.x] = wi - w;
    if (threadIdx.x == 31) tile_sum[blockIdx.x] = wi;
  }
  __syncthreads();
  unsigned run = s_warp[threadIdx.x >> 5] + incl - sum;
#pragma unroll
  for (int k = 0; k < kScanPer; k++) {
    if (first + k < n) hist[first + k] = run;
    run += v[k];
  }
}
__global__ void __launch_bounds__(kScanThreads) k_sort_scan_sums(unsigned* tile_sum, int tiles) {  // tiles <= 1024
  __shared__ unsigned s_warp[kScanThreads / 32];
  const unsigned v = (int)threadIdx.x < tiles ? tile_sum[threadIdx.x] : 0u;
  unsigned incl = v;
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
    if ((threadIdx.x & 31) >= off) incl += o;
  }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    const unsigned w = s_warp[threadIdx.x];
    unsigned wi = w;
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, wi, off);
      if (threadIdx.x >= off) wi += o;
    }
    s_warp[threadIdx.x] = wi - w;
  }
  __syncthreads();
  if ((int)threadIdx.x < tiles) tile_sum[threadIdx.x] = s_warp[threadIdx.x >> 5] + incl - v;
}
__global__ void k_sort_scatter_fine(const RenderArgs A, const int seg, const unsigned* __restrict__ tile_sum) {
  const unsigned n = seg == 0 ? (unsigned)A.npix * (unsigned)A.nsamp : A.q_count[kQHits0 + seg - 1];
  const unsigned miss = 1u << (3 * A.sort_bits);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned key = A.perm[i];
    unsigned dst;
    if (key == miss) {
      const unsigned m = __activemask();
      const int leader = __ffs(m) - 1, lane = threadIdx.x & 31;
      unsigned base = 0;
      if (lane == leader) base = atomicAdd(A.sort_hist + miss, (unsigned)__popc(m));
      base = __shfl_sync(m, base, leader);
      dst = base + __popc(m & ((1u << lane) - 1u));
    } else {
      dst = atomicAdd(A.sort_hist + key, 1u);
    }
    dst += tile_sum[key / kScanTile];
    const float4 o = A.ray_o[seg & 1][i], d = A.ray_d[seg & 1][i];
    A.sorted[2 * (size_t)dst] = A.hit[i];  // one full 32-byte sector per ray
    A.sorted[2 * (size_t)dst + 1] = make_float4(d.x, d.y, d.z, o.w);
  }
}
static void launch_sort_hits_fine(const RenderArgs& a, int seg, cudaStream_t st) {
  const unsigned n = (1u << (3 * a.sort_bits)) + 1;  // cells + the miss bucket
  const int tiles = (int)((n + kScanTile - 1) / kScanTile);
  unsigned* tile_sum = a.sort_hist + kSortFineMax + 2;
  cudaMemsetAsync(a.sort_hist, 0, (size_t)n * sizeof(unsigned), st);
  const int ctas = (a.num_sms > 0 ? a.num_sms : 148) * 8;
  k_sort_count_fine<<<ctas, 256, 0, st>>>(a, seg);
  k_sort_scan_tiles<<<tiles, kScanThreads, 0, st>>>(a.sort_hist, n, tile_sum);
  k_sort_scan_sums<<<1, kScanThreads, 0, st>>>(tile_sum, tiles);
  k_sort_scatter_fine<<<ctas, 256, 0, st>>>(a, seg, tile_sum);
}
void launch_sort_hits(const RenderArgs& a, int seg, cudaStream_t st) {
  if (a.sort_bits > 0) {
    launch_sort_hits_fine(a, seg, st);
    return;
  }
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_sort_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem);
    cudaFuncSetAttribute(k_sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem);
    configured = true;
  }
  cudaMemsetAsync(a.sort_hist, 0, (kSortBuckets + 2) * sizeof(unsigned), st);
  const int ctas = (a.num_sms > 0 ? a.num_sms : 148) * (kSortSmem <= 48 * 1024 ? 2 : 1);  // resident CTAs
  k_sort_count<<<ctas, kSortThreads, kSortSmem, st>>>(a, seg);
  k_sort_scan<<<1, 1024, 0, st>>>(a.sort_hist);
  k_sort_scatter<<<ctas, kSortThreads, kSortSmem, st>>>(a, seg);
}

// ----------------------------------------------------------------------------------------------
// k_shade: one thread per ray of the segment
// ----------------------------------------------------------------------------------------------
// Renderer.cpp:63-104
// A query in which two candidates tied (see kd_knearest_sorted): repeat it with the literal heap restatement,
// whose element moves are libstdc++'s, and hand the result back in the caller's arrays.  Rare by construction.
__device__ __noinline__ void knn_exact_redo(const DScene& S, float3 q, int k, unsigned long long* sc, int ks) {
  float hd[kMaxK];
  int hi[kMaxK];
  int kst[3 * kKdStack];
  KdHeap H{hd, hi, 1};
  unsigned long long visits = 0;
  kd_knearest(S, q, k, H, kst, 1, visits);
  for (int j = 0; j < k; j++) sc[j * ks] = kd_pack(hd[j], hi[j]);
}

// sc: this thread's k candidate slots (stride ks between slots), kst: its stack frames (stride kBlock)
RT_DI void knn_query(const DScene& S, float3 P, int k, int exact, unsigned long long* sc, int ks, int* kst,
                     unsigned long long& visits) {
  if (exact > 0)
    kd_knearest_exact(S, P, k, sc, ks, kst, kBlock, visits);
  else if (exact < 0)  // large k: libstdc++'s heap restated (log k moves per insertion, ties for free)
    kd_knearest_heap(S, P, k, sc, ks, kst, kBlock, visits);
  else if (kd_knearest_sorted(S, P, k, sc, ks, kst, kBlock, visits))
    knn_exact_redo(S, P, k, sc, ks);
}

// Renderer.cpp:88-103 from the gathered photons: r = distance of the k-th, avg = sum of the k incomeDirections in
// ascending distance; radiance = k (the sum of k ones, exact in binary32) / (pi r^2) / numPhotons * 100
RT_DI float4 gather_result(const DScene& S, int k, const unsigned long long* sc, int ks) {
  const float r = kd_dist_of(sc[(unsigned)(k - 1) * (unsigned)ks]);  // farthest of the k (ascending distance)
  float3 avg = f3(0.f, 0.f, 0.f);
  for (int j = 0; j < k; j++) avg = v_add(avg, f3(__ldg(S.kd_dir + kd_index_of(sc[(unsigned)j * (unsigned)ks]))));
  return make_float4(avg.x, avg.y, avg.z, r);
}
RT_DI float3 shade_photon_from(float4 g, float3 dir, float3 n, const DMaterial& m, int k, int num_photons) {
  const float r = g.w;
  float area = (float)__dmul_rn(__dmul_rn(3.141592653589793, (double)r), (double)r);
  float rad = __fmul_rn(__fdiv_rn(__fdiv_rn((float)k, area), (float)num_photons), 100.f);
  float3 bsdf = evaluate_color_response(m, n, v_norm(f3(g)), v_neg(dir));
  return v_scl(bsdf, rad);
}
RT_DI float3 shade_photon(const DScene& S, float3 dir, float3 n, float3 P, const DMaterial& m, int k, int num_photons,
                          int exact, unsigned long long* sc, int ks, int* kst, unsigned long long& visits) {
  knn_query(S, P, k, exact, sc, ks, kst, visits);
  return shade_photon_from(gather_result(S, k, sc, ks), dir, n, m, k, num_photons);
}
// dynamic shared memory of the kernels that run the k-NN: `frames` stack frames of 3 ints per thread, and -- up to
// k = kKnnSharedMaxK -- the k candidate (distance, index) pairs; beyond that the candidates live in global memory
size_t knn_smem_bytes(int k, int frames) {
  return (size_t)((k <= kKnnSharedMaxK ? 2 * k : 0) + 3 * frames) * kBlock * sizeof(int);
}

// ----------------------------------------------------------------------------------------------
// k_knn_gather: the k-nearest-photon queries of a segment's hit points as their own persistent kernel.
// ncu on the query inside k_shade (start of round 2; profiles/r2_cfg3_knn_seg0_sass_regions.txt is the shipped build): the traversal loop ran with 24 of 32 lanes (a warp's 32
// queries take 60-140 node visits each and the warp waits for the longest), the candidate insertion with 13 and the
// stack unwind with 7.  Here a query is a resumable state machine (KdQuery): every lane advances its own query by
// one node per round, a lane whose query is complete writes its result -- the sum of the k incomeDirections in
// ascending distance and the distance of the k-th, all the shading needs (Renderer.cpp:88-97) -- and the warp
// refills its idle lanes from the queue when fewer than kGatherRefillBelow are live, like k_trace does with rays.
// The queue is the sorted payload of the segment (hit points binned by Morton cell: neighbouring lanes walk the same
// part of the tree); slots whose ray missed are skipped.
// ----------------------------------------------------------------------------------------------
constexpr int kGatherRefillBelow = 24;
template <int POLICY, bool GSC>
__global__ void __launch_bounds__(kBlock, 6) k_knn_gather(const RenderArgs A, const int seg) {
  extern __shared__ unsigned long long s_knn[];
  const DScene& S = A.scene;
  const unsigned n = seg == 0 ? (unsigned)A.npix * (unsigned)A.nsamp : A.q_count[kQHits0 + seg - 1];
  unsigned* fetch = A.q_count + kQFetchAny0 + seg;  // photon mode casts no shadow rays: the any-hit cursor is free
  const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
  const int k = A.k;
  unsigned long long* sc;
  int cs;
  int* kst;
  if (GSC) {
    sc = A.knn_scratch + (size_t)blockIdx.x * kBlock + threadIdx.x;
    cs = A.knn_scratch_stride;
    kst = (int*)s_knn + threadIdx.x;
  } else {
    sc = s_knn + threadIdx.x;
    cs = kBlock;
    kst = (int*)(s_knn + k * kBlock) + threadIdx.x;
  }
  KdQuery<POLICY> Q;
  Q.nv = 0;
  unsigned slot = 0, n_knn = 0;
  unsigned long long n_visits = 0;
  bool live = false, exhausted = false;
  for (;;) {
    const unsigned need_mask = __ballot_sync(kFull, !live);
    if (!exhausted && need_mask) {
      const unsigned cnt = __popc(need_mask);
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(fetch, cnt);
      base = __shfl_sync(kFull, base, 0);
      if (base + cnt >= n) exhausted = true;
      if (!live) {
        const unsigned my = base + __popc(need_mask & lt_mask);
        if (my < n) {
          const float4 hr = (((A.sort_mask >> seg) & 1) && A.perm != nullptr) ? A.sorted[2 * (size_t)my] : A.hit[my];
          HitRec h;
          h.t = hr.x, h.u = hr.y, h.v = hr.z, h.gid = __float_as_int(hr.w);
          if (h.gid >= 0) {
            slot = my;
            Q.init(S, hit_point(S, h), k, sc, cs);
            live = true;
            n_knn++;
          }
        }
      }
    }
    if (__ballot_sync(kFull, live) == 0) {
      if (exhausted) break;
      continue;
    }
    for (;;) {
      if (live && !Q.step(S, k, sc, cs, kst, kBlock)) {  // this lane's query is complete
        if (POLICY == 0 && Q.tie) knn_exact_redo(S, Q.q, k, sc, cs);
        Q.finish(k, sc, cs);
        A.knn_out[slot] = gather_result(S, k, sc, cs);
        n_visits += Q.nv;
        live = false;
      }
      const unsigned live_mask = __ballot_sync(kFull, live);
      if (live_mask == 0) break;
      if (!exhausted && __popc(live_mask) < kGatherRefillBelow) break;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    n_knn += __shfl_xor_sync(kFull, n_knn, off);
    n_visits += __shfl_xor_sync(kFull, n_visits, off);
  }
  if (lane == 0) {
    if (n_knn) atomicAdd(A.counters + kCntKnn, (unsigned long long)n_knn);
    if (n_visits) atomicAdd(A.counters + kCntKdVisits, n_visits);
  }
}
template <int POLICY, bool GSC>
static int knn_gather_occupancy(size_t sm) {
  int occ = 0;
  cudaFuncSetAttribute(k_knn_gather<POLICY, GSC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_knn_gather<POLICY, GSC>, kBlock, sm);
  return occ < 1 ? 1 : occ;
}
int knn_gather_ctas_per_sm(int k, int kd_frames, int flavour) {
  const size_t sm = knn_smem_bytes(k, kd_frames);
  const bool g = k > kKnnSharedMaxK;
  if (flavour == 0) return g ? knn_gather_occupancy<0, true>(sm) : knn_gather_occupancy<0, false>(sm);
  return g ? knn_gather_occupancy<1, true>(sm) : knn_gather_occupancy<1, false>(sm);
}
void launch_knn_gather(const RenderArgs& a, int seg, cudaStream_t st) {
  const size_t sm = knn_smem_bytes(a.k, a.kd_frames);
  const int flavour = a.knn_exact < 0 ? 1 : 0;
  const bool g = a.k > kKnnSharedMaxK;
  int grid = (a.num_sms > 0 ? a.num_sms : 148) * knn_gather_ctas_per_sm(a.k, a.kd_frames, flavour);
  if (g && (size_t)grid * kBlock > (size_t)a.knn_scratch_stride) grid = a.knn_scratch_stride / kBlock;
  if (grid < 1) grid = 1;
  if (flavour == 0) {
    if (g) k_knn_gather<0, true><<<grid, kBlock, sm, st>>>(a, seg);
    else k_knn_gather<0, false><<<grid, kBlock, sm, st>>>(a, seg);
  } else {
    if (g) k_knn_gather<1, true><<<grid, kBlock, sm, st>>>(a, seg);
    else k_knn_gather<1, false><<<grid, kBlock, sm, st>>>(a, seg);
  }
}

// Photon k_shade: order the slots [tile, tile + R * kBlock) of a segment by the 30-bit Morton code of their hit points --
// bitonic sort of (code << 32 | tile-local index) in `sb` (the shared memory the queries use afterwards) -- and park
// the order in this tile's own piece of the perm array (idle between k_sort_scatter and the next k_sort_count).
// Slots past n sort last, misses just before them.
RT_DI void shade_tile_order(const RenderArgs& A, unsigned tile, unsigned n, int R, bool permuted, unsigned long long* sb) {
  const DScene& S = A.scene;
  const unsigned TILE = (unsigned)kBlock * (unsigned)R;
  for (int r = 0; r < R; r++) {
    const unsigned li = (unsigned)r * kBlock + threadIdx.x, slot = tile + li;
    unsigned key = 0xffffffffu;
    if (slot < n) {
      const float4 hr = permuted ? A.sorted[2 * (size_t)slot] : A.hit[slot];
      HitRec h;
      h.t = hr.x, h.u = hr.y, h.v = hr.z, h.gid = __float_as_int(hr.w);
      key = 0xfffffffeu;
      if (h.gid >= 0) {
        const float3 P = hit_point(S, h);
        const float sc10 = A.sort_key_scale;
        const int cx = min(max((int)((P.x - A.sort_lo.x) * A.sort_inv_cell.x * sc10), 0), 1023);
        const int cy = min(max((int)((P.y - A.sort_lo.y) * A.sort_inv_cell.y * sc10), 0), 1023);
        const int cz = min(max((int)((P.z - A.sort_lo.z) * A.sort_inv_cell.z * sc10), 0), 1023);
        key = morton3_10bit((unsigned)cx, (unsigned)cy, (unsigned)cz);
      }
    }
    sb[li] = ((unsigned long long)key << 32) | li;
  }
  __syncthreads();
  for (unsigned k2 = 2; k2 <= TILE; k2 <<= 1)
    for (unsigned j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
      for (unsigned q = threadIdx.x; q < TILE / 2; q += kBlock) {
        const unsigned i = 2 * q - (q & (j2 - 1));
        const unsigned long long x = sb[i], y = sb[i + j2];
        if ((x > y) == ((i & k2) == 0)) {
          sb[i] = y;
          sb[i + j2] = x;
        }
      }
      __syncthreads();
    }
  for (int r = 0; r < R; r++) {
    const unsigned li = (unsigned)r * kBlock + threadIdx.x;
    if (tile + li < n) A.perm[tile + li] = (unsigned)sb[li];  // low word: the tile-local index
  }
  __syncthreads();
}

// NLT: 3 = the stock three-light loop unrolled (Main.cpp:101-124), 0 = any light count (Renderer.cpp:49).
// GSC (PHOTON only): k-NN candidates in global memory (k > kKnnSharedMaxK).
template <int MODE, bool PHOTON, int NLT, bool GSC>
__global__ void __launch_bounds__(kBlock, PHOTON ? 6 : RT_SHADE_MINB) k_shade(const RenderArgs A, const int seg) {
  extern __shared__ unsigned long long s_knn[];  // PHOTON only: k 64-bit candidate rows, then 3*frames int rows
  const DScene& S = A.scene;
  const unsigned n = seg == 0 ? (unsigned)A.npix * (unsigned)A.nsamp : A.q_count[kQHits0 + seg - 1];
  const float4* qo_in = A.ray_o[seg & 1];
  const float4* qd_in = A.ray_d[seg & 1];
  float4* qo_out = A.ray_o[(seg + 1) & 1];
  float4* qd_out = A.ray_d[(seg + 1) & 1];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned nl = PHOTON ? 1u : (NLT > 0 ? (unsigned)NLT : (unsigned)S.num_lights);
  unsigned n_hit = 0, n_knn = 0;
  unsigned long long n_visits = 0;

  const bool permuted = ((A.sort_mask >> seg) & 1) && A.perm != nullptr;
  // PHOTON: a CTA takes tile_rounds * kBlock consecutive slots at a time and first orders them by the 30-bit Morton
  // code of their hit points (bitonic sort in the shared memory the queries use afterwards; the order is parked in
  // this tile's own piece of the perm array, idle at this point), so that the 32 queries of a warp start a few
  // millimetres apart and walk the photon tree in lock-step.  The global binning (k_sort_*) makes consecutive slots
  // neighbours at the scale of its cells; this refines it for ~400 instructions per query.  The order only decides
  // which lane runs which query: results do not depend on it.
  const int R = PHOTON ? max(A.tile_rounds, 1) : 1;
  const unsigned TILE = (unsigned)kBlock * (unsigned)R;
  const bool tile_sort = PHOTON && A.tile_rounds > 1;
  // tiles are handed out by a cursor (the any-hit fetch counter, idle in photon mode): a tile of 1024 queries is ~4 % of a
  // CTA's share of a segment, and static round-robin left the SMs waiting for the CTAs with the expensive tiles
  __shared__ unsigned s_tile;
  unsigned* tile_cursor = A.q_count + kQFetchAny0 + seg;
  for (unsigned tile = blockIdx.x * TILE;; tile += gridDim.x * TILE) {
    if (PHOTON && tile_sort) {
      if (threadIdx.x == 0) s_tile = atomicAdd(tile_cursor, 1u);
      __syncthreads();
      tile = s_tile * TILE;  // (the previous tile's closing barrier keeps s_tile stable until everyone has read it)
    }
    if (tile >= n) break;
    if (PHOTON && tile_sort) shade_tile_order(A, tile, n, R, permuted, s_knn);
    for (int r = 0; r < R; r++) {
      unsigned li = (unsigned)r * kBlock + threadIdx.x;
      // (written above by this very thread; slots past n sort last, so the first n - tile entries are the valid ones)
      if (PHOTON && tile_sort && tile + li < n) li = A.perm[tile + li];
      const unsigned slot = tile + li;
      bool found = false;
      unsigned p = 0;
      float3 d = f3(0, 0, 0);
      HitRec h;
      if (slot < n) {
        float4 b, hr;
        if (permuted) {  // sorted payload written by k_sort_scatter: direction + path id, hit record
          hr = A.sorted[2 * (size_t)slot];
          b = A.sorted[2 * (size_t)slot + 1];
          p = (unsigned)__float_as_int(b.w);
        } else {
          b = qd_in[slot];
          hr = A.hit[slot];
          p = (unsigned)__float_as_int(qo_in[slot].w);
        }
        d = f3(b);
        h.t = hr.x;
        h.u = hr.y;
        h.v = hr.z;
        h.gid = __float_as_int(hr.w);
        found = h.gid >= 0;
      }
      // compacted index j of this hit (warp-aggregated)
      const unsigned mask = __ballot_sync(kFull, found);
      unsigned j = 0;
      if (mask) {
        unsigned j0 = 0;
        if (lane == (unsigned)(__ffs(mask) - 1)) j0 = atomicAdd(A.q_count + kQHits0 + seg, (unsigned)__popc(mask));
        j0 = __shfl_sync(kFull, j0, __ffs(mask) - 1);
        j = j0 + __popc(mask & ((1u << lane) - 1u));
      }
      if (slot < n && !found) {
        // Renderer.cpp:154-160: a miss ends the path and contributes Vec3f(0) from this segment on.  The segments' colours
        // live in their own arrays and are added -- c0 + (c1 + c2), Renderer.cpp:168 -- and clamped by k_resolve, so a
        // path that ends here only zeroes the terms it will never write (no read-modify-write of earlier ones).
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        if (seg == 0) A.col0[p] = zero;  // w = 0: posIntersectionFound = false
        if (MODE == 1) {
          if (seg <= 1) A.col1[p] = zero;
          A.col2[p] = zero;
        }
      }
      if (found) {
        n_hit++;
        int pixel, sample;
        Rng g;
        // words consumed before this segment: 4 (jitter), then per earlier segment 2 per light + 4 (hemisphere);
        // the photon gather draws nothing (Renderer.cpp:63-104)
        g.init(path_stream_key(A, p, pixel, sample),
               4u + (PHOTON ? 4u : 2u * (unsigned)S.num_lights + 4u) * (unsigned)seg);
        float3 nrm, P, tp0, te1, te2;
        int mesh;
        hit_geometry(S, h, nrm, P, mesh, tp0, te1, te2);
        const DMaterial m = S.mats[mesh];
        if (!PHOTON) A.hit_path[j] = (int)p;
        // PHOTON: no shadow rays, so nothing waits for an any-hit answer: the segment's colour goes straight to the
        // path's slot -- what k_combine would do with it one launch later (`0 + contribution`, Renderer.cpp:44,59;
        // seg 0 of -m 0 clamped, Renderer.cpp:254), without the trip through contrib[] and hit_path[]
        auto write_photon_colour = [&](float3 c) {
          c = v_add(f3(0.f, 0.f, 0.f), c);
          if (seg == 0) {
            const float3 out = MODE == 0 ? normalize_color(c) : c;
            A.col0[p] = make_float4(out.x, out.y, out.z, 1.f);
          } else if (seg == 1) {
            A.col1[p] = make_float4(c.x, c.y, c.z, 0.f);
          } else {
            A.col2[p] = make_float4(c.x, c.y, c.z, 0.f);
          }
        };
        if (PHOTON && A.knn_out != nullptr) {  // gathered by k_knn_gather, per ray slot of the segment
          write_photon_colour(shade_photon_from(A.knn_out[slot], d, nrm, m, A.k, A.num_photons));
        } else if (PHOTON) {
          n_knn++;
          unsigned long long* sc;
          int ks;
          int* kst;
          if (GSC) {
            sc = A.knn_scratch + (size_t)blockIdx.x * kBlock + threadIdx.x;
            ks = A.knn_scratch_stride;
            kst = (int*)s_knn + threadIdx.x;
          } else {
            sc = s_knn + threadIdx.x;
            ks = kBlock;
            kst = (int*)(s_knn + A.k * kBlock) + threadIdx.x;
          }
          write_photon_colour(shade_photon(S, d, nrm, P, m, A.k, A.num_photons, A.knn_exact, sc, ks, kst, n_visits));
        } else {
          // Renderer.cpp:49-60: per light 2 uniforms, the shadow ray, and (eagerly) radiance * bsdf
          const BsdfFrame bf = bsdf_frame(m, nrm, v_neg(d));
          A.hit_p[j] = make_float4(P.x, P.y, P.z, 0.f);
  #pragma unroll
          for (int l = 0; l < (NLT > 0 ? NLT : (int)nl); l++) {
            const unsigned s = shadow_slot(j, (unsigned)l, nl);
            const DLight& L = light_at(S, l);
            float3 to_light = v_sub(light_rand_area_position(L, g), P);
            float3 c = v_mul(light_evaluate(L, P), evaluate_color_response(m, bf, to_light));
            // Any-hit is an OR over all triangles (Renderer.cpp:52-55, RayTracer.h:40), so testing the triangle the ray
            // starts on first cannot change the answer: 12-15 % of the shadow rays end here (shadow acne is part of
            // the reference image), at this kernel's ~97 % lane use instead of the traversal's ~65 %.
            bool self = false;
            if (A.own_tri) {
              float tu, tv, tt;
              self = mt_intersect(P, to_light, tp0, te1, te2, tu, tv, tt) && tt > 0.f && tt < FLT_MAX;
            }
            A.sh_d[s] = make_float4(to_light.x, to_light.y, to_light.z, self ? 1.f : 0.f);
            A.contrib[s] = make_float4(c.x, c.y, c.z, 0.f);
            if (self) A.occ[s] = 1;
          }
        }
        if (MODE == 1 && seg < 2) {  // Renderer.cpp:164-166: bounce
          float3 nd = hsphere_uniform_sample(g, nrm);
          qo_out[j] = make_float4(P.x, P.y, P.z, __int_as_float((int)p));
          qd_out[j] = make_float4(nd.x, nd.y, nd.z, 0.f);
        }
      }
    }
    if (PHOTON && tile_sort) __syncthreads();  // the next tile's sort reuses the queries' shared memory
  }
  for (int off = 16; off > 0; off >>= 1) {
    n_hit += __shfl_xor_sync(kFull, n_hit, off);
    n_knn += __shfl_xor_sync(kFull, n_knn, off);
    n_visits += __shfl_xor_sync(kFull, n_visits, off);
  }
  if (lane == 0) {
    if (n_hit && !PHOTON) atomicAdd(A.counters + kCntShadow, (unsigned long long)n_hit * A.scene.num_lights);
    if (n_knn) atomicAdd(A.counters + kCntKnn, (unsigned long long)n_knn);
    if (n_visits) atomicAdd(A.counters + kCntKdVisits, n_visits);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(A.counters + kCntNearest, (unsigned long long)n);
}

template <int MODE, bool GSC>
static int shade_photon_occupancy(size_t sm) {
  int occ = 0;
  cudaFuncSetAttribute(k_shade<MODE, true, 0, GSC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_shade<MODE, true, 0, GSC>, kBlock, sm);
  return occ < 1 ? 1 : occ;
}
int shade_photon_ctas_per_sm(int mode, int k, int kd_frames) {
  const size_t sm = knn_smem_bytes(k, kd_frames);
  const bool g = k > kKnnSharedMaxK;
  if (mode == 0) return g ? shade_photon_occupancy<0, true>(sm) : shade_photon_occupancy<0, false>(sm);
  return g ? shade_photon_occupancy<1, true>(sm) : shade_photon_occupancy<1, false>(sm);
}

void launch_shade(const RenderArgs& a, int seg, int grid, cudaStream_t st) {
  if (a.photon) {
    const size_t sm = a.knn_out ? 0 : knn_smem_bytes(a.k, a.kd_frames);
    // grid-stride kernel: launch exactly what is resident (the k-NN shared memory allows 6 CTAs/SM at k = 10, 2 at
    // k = 50; 8 per SM left a third of the CTAs for a second, mostly empty wave -- 26 % warps active in ncu)
    const int occ = a.knn_out ? 8 : shade_photon_ctas_per_sm(a.mode, a.k, a.kd_frames);
    if (a.num_sms > 0) grid = grid < a.num_sms * occ ? grid : a.num_sms * occ;
    const bool g = a.k > kKnnSharedMaxK;
    if (g && (size_t)grid * kBlock > (size_t)a.knn_scratch_stride) grid = a.knn_scratch_stride / kBlock;
    // a tile is a CTA's unit of work: keep at least four of them per resident CTA, or the tail (and, for a single
    // sample per pixel, half of the SMs) idles -- -m 0 -N 1 -p 500000 -k 50 went from 3.7 to 9.0 ms with 1024-slot tiles
    RenderArgs b = a;
    const long long slots = (long long)a.npix * a.nsamp;
    if (b.tile_rounds < 0) b.tile_rounds = -b.tile_rounds;  // forced (tests: small frames must take the tiled path too)
    else
      while (b.tile_rounds > 1 && slots < 4LL * grid * kBlock * b.tile_rounds) b.tile_rounds /= 2;
    if (a.mode == 0) {
      if (g) k_shade<0, true, 0, true><<<grid, kBlock, sm, st>>>(b, seg);
      else k_shade<0, true, 0, false><<<grid, kBlock, sm, st>>>(b, seg);
    } else {
      if (g) k_shade<1, true, 0, true><<<grid, kBlock, sm, st>>>(b, seg);
      else k_shade<1, true, 0, false><<<grid, kBlock, sm, st>>>(b, seg);
    }
    return;
  }
  const bool three = a.scene.num_lights == 3;
  {  // grid-stride kernel: launch exactly what is resident (RT_SHADE_MINB CTAs per SM unless the registers say otherwise)
    static int occ[2][2] = {{0, 0}, {0, 0}};
    int& o = occ[a.mode == 1][three];
    if (o == 0) {
      if (a.mode == 0 && three) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_shade<0, false, 3, false>, kBlock, 0);
      if (a.mode == 0 && !three) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_shade<0, false, 0, false>, kBlock, 0);
      if (a.mode == 1 && three) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_shade<1, false, 3, false>, kBlock, 0);
      if (a.mode == 1 && !three) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_shade<1, false, 0, false>, kBlock, 0);
      if (o < 1) o = 1;
    }
    const long long blocks = ((long long)a.npix * a.nsamp + kBlock - 1) / kBlock;
    grid = (int)std::max<long long>(1, std::min<long long>(blocks, (long long)(a.num_sms > 0 ? a.num_sms : 148) * o));
  }
  if (a.mode == 0) {
    if (three) k_shade<0, false, 3, false><<<grid, kBlock, 0, st>>>(a, seg);
    else k_shade<0, false, 0, false><<<grid, kBlock, 0, st>>>(a, seg);
  } else {
    if (three) k_shade<1, false, 3, false><<<grid, kBlock, 0, st>>>(a, seg);
    else k_shade<1, false, 0, false><<<grid, kBlock, 0, st>>>(a, seg);
  }
}

// ----------------------------------------------------------------------------------------------
// k_combine: colorResponse = ((0 + a0) + a1) + a2 ... over the unoccluded lights in scene order
// (Renderer.cpp:44,49-60), then the path sum c0 + (c1 + c2) and the per-sample clamp (Renderer.cpp:168,254)
// ----------------------------------------------------------------------------------------------
template <int NLT>
__global__ void __launch_bounds__(256) k_combine(const RenderArgs A, const int seg) {
  const unsigned n = A.q_count[kQHits0 + seg];
  const unsigned nl = NLT > 0 ? (unsigned)NLT : (unsigned)A.nl;
  for (unsigned j = blockIdx.x * 256 + threadIdx.x; j < n; j += gridDim.x * 256) {
    float3 c = f3(0.f, 0.f, 0.f);  // (the photon gather writes its colour in k_shade: this kernel does not run for it)
#pragma unroll
    for (int l = 0; l < (NLT > 0 ? NLT : (int)nl); l++) {
      const unsigned s = shadow_slot(j, (unsigned)l, nl);
      if (A.occ[s] == 0) c = v_add(c, f3(A.contrib[s]));
    }
    const unsigned p = (unsigned)A.hit_path[j];
    // one write per hit and no read: hits of the bounce segments arrive in Morton order, so p is scattered; the path
    // sum and the per-sample clamp (Renderer.cpp:168,254) happen in k_resolve, which reads the three arrays coalesced
    // (with the read-modify-write of col0 here k_combine took 2.26 ms per cfg2 frame, 0.80 ms when the hits were
    // left unsorted)
    if (seg == 0) {
      float3 out = A.mode == 0 ? normalize_color(c) : c;
      A.col0[p] = make_float4(out.x, out.y, out.z, 1.f);
    } else if (seg == 1) {
      A.col1[p] = make_float4(c.x, c.y, c.z, 0.f);
    } else {
      A.col2[p] = make_float4(c.x, c.y, c.z, 0.f);
    }
  }
}
void launch_combine(const RenderArgs& a, int seg, int grid, cudaStream_t st) {
  if (a.nl == 3)
    k_combine<3><<<grid, 256, 0, st>>>(a, seg);
  else
    k_combine<0><<<grid, 256, 0, st>>>(a, seg);
}

// ----------------------------------------------------------------------------------------------
// ordered accumulation: updateImage(x,y) += colorResponse for samples in index order
// ----------------------------------------------------------------------------------------------
// path colour of one sample: -m 0: col0 holds the clamped colour already; -m 1: colorResponse = clamp(c0 + (c1 + c2))
// (Renderer.cpp:168: color + calculateColorPath(...), recursion depth 3; :254 normalizeColor)
RT_DI float4 path_colour(const float4* __restrict__ col0, const float4* __restrict__ col1, const float4* __restrict__ col2,
                         size_t p, int mode) {
  float4 c = col0[p];
  if (mode == 1) {
    const float3 out = normalize_color(v_add(f3(c), v_add(f3(col1[p]), f3(col2[p]))));
    c = make_float4(out.x, out.y, out.z, c.w);
  }
  return c;
}
// the same written back into col0 (rt_render_samples hands the per-sample colours to the caller)
__global__ void k_finalize_paths(float4* col0, const float4* __restrict__ col1, const float4* __restrict__ col2,
                                 long long n, int mode) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) col0[p] = path_colour(col0, col1, col2, (size_t)p, mode);
}
void launch_finalize_paths(float4* col0, const float4* col1, const float4* col2, long long n, int mode, cudaStream_t st) {
  if (n > 0 && mode == 1) k_finalize_paths<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(col0, col1, col2, n, mode);
}
__global__ void k_resolve(const float4* __restrict__ col0, const float4* __restrict__ col1,
                          const float4* __restrict__ col2, int mode, int npix, int nsamp, float4* acc_rgb, int* acc_cnt) {
  int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= npix) return;
  float4 acc = acc_rgb[pl];
  int cnt = acc_cnt[pl];
  for (int s = 0; s < nsamp; s++) {
    float4 c = path_colour(col0, col1, col2, (size_t)s * npix + pl, mode);
    acc.x = __fadd_rn(acc.x, c.x);
    acc.y = __fadd_rn(acc.y, c.y);
    acc.z = __fadd_rn(acc.z, c.z);
    cnt += (c.w != 0.f) ? 1 : 0;  // Renderer.cpp:255-257
  }
  acc_rgb[pl] = acc;
  acc_cnt[pl] = cnt;
}
void launch_resolve(const float4* col0, const float4* col1, const float4* col2, int mode, int npix, int nsamp,
                    float4* acc_rgb, int* acc_cnt, cudaStream_t st) {
  k_resolve<<<(npix + 255) / 256, 256, 0, st>>>(col0, col1, col2, mode, npix, nsamp, acc_rgb, acc_cnt);
}

__global__ void k_scatter(const float4* __restrict__ acc_rgb, const int* __restrict__ acc_cnt,
                          const int* __restrict__ pix_map, int npix, float* out_rgb, int* out_cnt) {
  int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= npix) return;
  int pixel = pix_map[pl];
  float4 a = acc_rgb[pl];
  out_rgb[3 * (size_t)pixel] = a.x;
  out_rgb[3 * (size_t)pixel + 1] = a.y;
  out_rgb[3 * (size_t)pixel + 2] = a.z;
  out_cnt[pixel] = acc_cnt[pl];
}
void launch_scatter(const float4* acc_rgb, const int* acc_cnt, const int* pix_map, int npix, float* out_rgb,
                    int* out_cnt, cudaStream_t st) {
  k_scatter<<<(npix + 255) / 256, 256, 0, st>>>(acc_rgb, acc_cnt, pix_map, npix, out_rgb, out_cnt);
}

__global__ void k_scatter_packed(const float4* __restrict__ acc_rgb, const int* __restrict__ acc_cnt,
                                 const int* __restrict__ pix_map, int npix, float4* out) {
  int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= npix) return;
  const float4 a = acc_rgb[pl];
  out[pix_map[pl]] = make_float4(a.x, a.y, a.z, (float)acc_cnt[pl]);  // counts < 2^24 are exact in binary32
}
void launch_scatter_packed(const float4* acc_rgb, const int* acc_cnt, const int* pix_map, int npix, float4* out,
                           cudaStream_t st) {
  k_scatter_packed<<<(npix + 255) / 256, 256, 0, st>>>(acc_rgb, acc_cnt, pix_map, npix, out);
}
// Renderer.cpp:262-265 on the packed frame: the counter arrives as a float holding an exact integer
__global__ void k_composite_packed(const float4* __restrict__ sum_rgbn, long long npx, int num_rays, float* rgb_inout) {
  const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= npx) return;
  const float4 a = sum_rgbn[px];
  const float fn = (float)num_rays, miss = (float)(num_rays - (int)a.w);
  const float s[3] = {a.x, a.y, a.z};
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    const float bg = rgb_inout[3 * px + ch];
    rgb_inout[3 * px + ch] = __fadd_rn(__fdiv_rn(s[ch], fn), __fdiv_rn(__fmul_rn(bg, miss), fn));
  }
}
void launch_composite_packed(const float4* sum_rgbn, long long npx, int num_rays, float* rgb_inout, cudaStream_t st) {
  k_composite_packed<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(sum_rgbn, npx, num_rays, rgb_inout);
}

// Renderer.cpp:262-265 after the last pass (i + 1 == N), on the pixels this shard owns:
//   saveImage = updateImage / float(N) + image * (N - counter) / float(N)
__global__ void k_composite(const float4* __restrict__ acc_rgb, const int* __restrict__ acc_cnt,
                            const int* __restrict__ pix_map, int npix, int num_rays, const float* background,
                            float* out) {
  int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= npix) return;
  const size_t pixel = (size_t)pix_map[pl];
  const float4 a = acc_rgb[pl];
  const float fn = (float)num_rays, miss = (float)(num_rays - acc_cnt[pl]);
  const float s[3] = {a.x, a.y, a.z};
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    const float bg = background[3 * pixel + ch];
    out[3 * pixel + ch] = __fadd_rn(__fdiv_rn(s[ch], fn), __fdiv_rn(__fmul_rn(bg, miss), fn));
  }
}
// the same composite on full-frame sums/counters (after the multi-GPU reduce): rgb_inout holds the background
__global__ void k_composite_frame(const float* __restrict__ sum_rgb, const int* __restrict__ counter, long long npx,
                                  int num_rays, float* rgb_inout) {
  const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= npx) return;
  const float fn = (float)num_rays, miss = (float)(num_rays - counter[px]);
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    const float bg = rgb_inout[3 * px + ch];
    rgb_inout[3 * px + ch] = __fadd_rn(__fdiv_rn(sum_rgb[3 * px + ch], fn), __fdiv_rn(__fmul_rn(bg, miss), fn));
  }
}
void launch_composite_frame(const float* sum_rgb, const int* counter, long long npx, int num_rays, float* rgb_inout,
                            cudaStream_t st) {
  k_composite_frame<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(sum_rgb, counter, npx, num_rays, rgb_inout);
}
// background and out may be the same buffer (rt_render composites in place)
void launch_composite(const float4* acc_rgb, const int* acc_cnt, const int* pix_map, int npix, int num_rays,
                      const float* background, float* out, cudaStream_t st) {
  k_composite<<<(npix + 255) / 256, 256, 0, st>>>(acc_rgb, acc_cnt, pix_map, npix, num_rays, background, out);
}

// ----------------------------------------------------------------------------------------------
// parity hooks
// ----------------------------------------------------------------------------------------------
// RayTracer::hsphereUniformSample around n normals; item i draws from stream (seed, domain, index0 + i), word 0 on
__global__ void k_hsphere(uint64_t seed_mixed, uint64_t domain, uint64_t index0, const float* __restrict__ normals,
                          long long n, float* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng g;
  g.init(stream_key(seed_mixed, domain, index0 + (uint64_t)i), 0);
  const float3 r = hsphere_uniform_sample(g, f3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]));
  out[3 * i] = r.x;
  out[3 * i + 1] = r.y;
  out[3 * i + 2] = r.z;
}
void launch_hsphere(uint64_t seed_mixed, uint64_t domain, uint64_t index0, const float* normals, long long n, float* out,
                    cudaStream_t st) {
  k_hsphere<<<(int)((n + 127) / 128), 128, 0, st>>>(seed_mixed, domain, index0, normals, n, out);
}

__global__ void k_bsdf(DMaterial m, const float* __restrict__ in, long long n, float* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* a = in + 9 * i;
  float3 r = evaluate_color_response(m, f3(a[0], a[1], a[2]), f3(a[3], a[4], a[5]), f3(a[6], a[7], a[8]));
  out[3 * i] = r.x;
  out[3 * i + 1] = r.y;
  out[3 * i + 2] = r.z;
}
void launch_bsdf(DMaterial m, const float* n_wi_wo, long long n, float* rgb, cudaStream_t st) {
  k_bsdf<<<(int)((n + 127) / 128), 128, 0, st>>>(m, n_wi_wo, n, rgb);
}

template <bool GSC>
__global__ void __launch_bounds__(kBlock) k_knn(const DScene S, const float* __restrict__ q3, long long n, int k,
                                                int exact, int* node_index, unsigned long long* counters,
                                                unsigned long long* scratch) {
  extern __shared__ unsigned long long s_knn[];
  unsigned long long* sc;
  int cs;
  int* kst;
  if (GSC) {  // k > kKnnSharedMaxK: candidates in global memory, [slot][resident thread]
    sc = scratch + (size_t)blockIdx.x * kBlock + threadIdx.x;
    cs = (int)(gridDim.x * kBlock);
    kst = (int*)s_knn + threadIdx.x;
  } else {
    sc = s_knn + threadIdx.x;
    cs = kBlock;
    kst = (int*)(s_knn + k * kBlock) + threadIdx.x;
  }
  unsigned long long visits = 0, queries = 0;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    const float3 q = f3(q3[3 * i], q3[3 * i + 1], q3[3 * i + 2]);
    knn_query(S, q, k, exact, sc, cs, kst, visits);
    for (int j = 0; j < k; j++) node_index[i * k + j] = kd_index_of(sc[(unsigned)j * (unsigned)cs]);
    queries++;
  }
  if (queries) {
    atomicAdd(counters + kCntKdVisits, visits);
    atomicAdd(counters + kCntKnn, queries);
  }
}
int knn_ctas_per_sm(int k, int kd_frames) {
  const size_t sm = knn_smem_bytes(k, kd_frames);
  int occ = 0;
  if (k > kKnnSharedMaxK) {
    cudaFuncSetAttribute(k_knn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_knn<true>, kBlock, sm);
  } else {
    cudaFuncSetAttribute(k_knn<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_knn<false>, kBlock, sm);
  }
  return occ < 1 ? 1 : occ;
}
void launch_knn(const DScene& s, const float* q3, long long n, int k, int kd_frames, int exact, int* node_index,
                unsigned long long* counters, unsigned long long* scratch, int grid, cudaStream_t st) {
  const size_t sm = knn_smem_bytes(k, kd_frames);
  if (grid < 1) grid = 1;
  if (k > kKnnSharedMaxK)
    k_knn<true><<<grid, kBlock, sm, st>>>(s, q3, n, k, exact, node_index, counters, scratch);
  else
    k_knn<false><<<grid, kBlock, sm, st>>>(s, q3, n, k, exact, node_index, counters, scratch);
}

// ----------------------------------------------------------------------------------------------
// k_emit: photon paths (PhotonMap.h:19-44 emission, :92-155 random walk).
// out_a = (position, weight), out_b = (incomeDirection, status bits): bit0 stored,
// bits 8.. = 1 + depth of the Russian-roulette kill (0: not counted in the histogram).
//
// A path is 1-20 bounces and most end after one or two (PhotonMap.h:144-150: Russian roulette), so "one thread walks
// one path" left 3.8 of 32 lanes busy (profiles/r2_cfg4_emit_full.csv, first capture): a warp ran as long as its
// longest path.  Here the unit of a loop iteration is ONE BOUNCE: every lane advances its own path by one segment
// (trace + shade + roulette), and a lane whose path has ended takes the next path from a global cursor (warp-
// aggregated), so the warp stays full until the cursor runs out.  Results do not depend on which lane walks which
// path: every path owns its random stream and its output slot q.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_emit(const DScene S, uint64_t seed_mixed, int per_light, float light_pdf,
                                                 int first_path, int npaths, int brute, float4* out_a, float4* out_b,
                                                 unsigned long long* counters, unsigned* cursor) {
  __shared__ int s_stack[kStackDepth * kBlock];
  int* stack = s_stack + threadIdx.x;
  const unsigned total = (unsigned)S.num_lights * (unsigned)npaths;
  const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
  unsigned n_rays = 0;
  // the path this lane is walking
  bool alive = false, exhausted = false;
  unsigned q = 0;
  Rng g;
  float3 o = f3(0, 0, 0), d = f3(0, 0, 0), ppos = f3(0, 0, 0), pdir = f3(0, 0, 0);
  float weight = 0.f;
  int depth = 0;
  for (;;) {
    const unsigned need = __ballot_sync(kFull, !alive);
    if (!exhausted && need) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(cursor, (unsigned)__popc(need));
      base = __shfl_sync(kFull, base, 0);
      if (base + __popc(need) >= total) exhausted = true;
      if (!alive) {
        q = base + __popc(need & lt_mask);
        if (q < total) {  // PhotonMap.h:24-37: a new path leaves light li
          const int li = (int)(q / (unsigned)npaths);
          const int path = first_path + (int)(q - (unsigned)li * (unsigned)npaths);
          const DLight& L = light_at(S, li);
          g.init(stream_key(seed_mixed, kDomainPhoton, (uint64_t)li * (uint64_t)per_light + (uint64_t)path), 0);
          o = light_rand_area_position(L, g);
          d = hsphere_uniform_sample(g, L.normal);
          const float pdf0 = v_dot(v_norm(d), v_norm(L.normal));
          weight = __fdiv_rn(light_radiance(L, o), __fmul_rn(pdf0, light_pdf));
          ppos = pdir = f3(0.f, 0.f, 0.f);
          depth = 0;
          alive = true;
        }
      }
    }
    if (__ballot_sync(kFull, alive) == 0) {
      if (exhausted) break;
      continue;
    }
    if (alive) {  // one segment of calculatePhotonPath at `depth` (< 20)
      HitRec h;
      n_rays++;
      const bool f = brute ? brute_trace<false>(S, o, d, h) : bvh_traverse<false>(S, o, d, stack, kBlock, h);
      bool stored = false, done = false;
      int hist = 0;
      if (!f) {  // PhotonMap.h:109-112: the photon left the scene; it is stored where it last hit
        stored = depth != 0;
        done = true;
      } else {
        float3 nrm, P;
        int mesh;
        hit_geometry(S, h, nrm, P, mesh);
        const DMaterial m = S.mats[mesh];
        ppos = P;
        pdir = v_neg(d);
        const float3 rd = hsphere_uniform_sample(g, nrm);
        const float3 refl = v_sub(d, v_scl(nrm, __fmul_rn(2.f, v_dot(d, nrm))));
        const float bsdf = v_len(evaluate_color_response(m, nrm, d, rd));
        const float pdf = __fdiv_rn(__fadd_rn(v_dot(v_norm(rd), v_norm(refl)), 1.f), 2.f);
        weight = __fmul_rn(weight, __fdiv_rn(bsdf, pdf));
        const float cont = fminf(weight, 1.f);
        if (g.uniform_f(0.f, 1.f) > cont) {  // PhotonMap.h:144-150, then :94-97 on entry at depth + 1
          stored = true;
          hist = depth + 1;
          done = true;
        } else {
          weight = __fdiv_rn(weight, cont);
          o = P;
          d = rd;
          depth++;
          done = depth >= 20;  // PhotonMap.h:98: dropped, not stored
        }
      }
      if (done) {
        out_a[q] = make_float4(ppos.x, ppos.y, ppos.z, weight);
        out_b[q] = make_float4(pdir.x, pdir.y, pdir.z, __int_as_float((stored ? 1 : 0) | (hist << 8)));
        alive = false;
      }
    }
  }
  for (int off = 16; off > 0; off >>= 1) n_rays += __shfl_xor_sync(kFull, n_rays, off);
  if (lane == 0 && n_rays) atomicAdd(counters + kCntPhotonRays, (unsigned long long)n_rays);
}
void launch_emit(const DScene& s, uint64_t seed_mixed, int per_light, float light_pdf, int first_path, int npaths,
                 int brute, float4* out_a, float4* out_b, unsigned long long* counters, unsigned* cursor, int num_sms,
                 cudaStream_t st) {
  long long total = (long long)s.num_lights * npaths;
  long long blocks = (total + kBlock - 1) / kBlock;
  // persistent: a few CTAs per SM share the cursor (20 KB of traversal stack each)
  const long long resident = (long long)(num_sms > 0 ? num_sms : 148) * 8;
  if (blocks > resident) blocks = resident;
  if (blocks < 1) blocks = 1;
  cudaMemsetAsync(cursor, 0, sizeof(unsigned), st);
  k_emit<<<(int)blocks, kBlock, 0, st>>>(s, seed_mixed, per_light, light_pdf, first_path, npaths, brute, out_a, out_b,
                                         counters, cursor);
}

// ----------------------------------------------------------------------------------------------
// Device-resident photon list (multi-GPU path): the particles k_emit stored are compacted in
// (light, path) order -- the order PhotonMap.h:94-96,109-111 appends in -- without leaving the GPU, so that the
// NCCL all-gather can read them where they are.  Three small kernels: per-block counts (+ per-light counts and the
// Russian-roulette depth histogram), a one-block scan of the block counts, the ordered write.
// ----------------------------------------------------------------------------------------------
constexpr int kCompactBlock = 1024;
__global__ void __launch_bounds__(kCompactBlock) k_photon_count(const float4* __restrict__ out_b, long long total,
                                                                int npaths, unsigned* block_count,
                                                                unsigned long long* light_count, unsigned* hist20) {
  __shared__ unsigned s_warp[kCompactBlock / 32];
  const long long q = (long long)blockIdx.x * kCompactBlock + threadIdx.x;
  int status = 0;
  if (q < total) status = __float_as_int(out_b[q].w);
  const bool stored = (status & 1) != 0;
  const int hist = status >> 8;
  if (hist > 0 && hist <= 20) atomicAdd(hist20 + hist - 1, 1u);
  const unsigned m = __ballot_sync(kFull, stored);
  if (stored) {  // per-light counts: one atomic per warp when the whole warp sits in one light (almost always)
    const int li = (int)(q / npaths);
    const unsigned peers = __match_any_sync(m, li);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(light_count + li, (unsigned long long)__popc(peers));
  }
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned v = s_warp[threadIdx.x];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    if (threadIdx.x == 0) block_count[blockIdx.x] = v;
  }
}
// exclusive scan of nb counters in place by one CTA; total -> counts[nb]
__global__ void __launch_bounds__(1024) k_scan_u32(unsigned* counts, int nb) {
  __shared__ unsigned s[1024];
  unsigned carry = 0;
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned v = i < nb ? counts[i] : 0u;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      const unsigned add = threadIdx.x >= off ? s[threadIdx.x - off] : 0u;
      __syncthreads();
      s[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < nb) counts[i] = carry + s[threadIdx.x] - v;
    const unsigned tot = s[1023];
    __syncthreads();
    carry += tot;
  }
  if (threadIdx.x == 0) counts[nb] = carry;
}
__global__ void __launch_bounds__(kCompactBlock) k_photon_compact(const float4* __restrict__ out_a,
                                                                  const float4* __restrict__ out_b, long long total,
                                                                  const unsigned* __restrict__ block_offset, float* out7,
                                                                  long long capacity) {
  __shared__ unsigned s_warp[kCompactBlock / 32];
  const long long q = (long long)blockIdx.x * kCompactBlock + threadIdx.x;
  float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
  bool stored = false;
  if (q < total) {
    b = out_b[q];
    stored = (__float_as_int(b.w) & 1) != 0;
    if (stored) a = out_a[q];
  }
  const unsigned m = __ballot_sync(kFull, stored);
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  if (threadIdx.x < 32) {  // exclusive scan of the 32 warp counts
    const unsigned v = s_warp[threadIdx.x];
    unsigned inc = v;
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned o = __shfl_up_sync(kFull, inc, off);
      if (threadIdx.x >= (unsigned)off) inc += o;
    }
    s_warp[threadIdx.x] = inc - v;
  }
  __syncthreads();
  if (stored) {
    const long long dst = (long long)block_offset[blockIdx.x] + s_warp[warp] + __popc(m & ((1u << lane) - 1u));
    if (dst < capacity) {
      float* o = out7 + 7 * dst;  // Particle{position, incomeDirection, weight}, Particle.h:33-35
      o[0] = a.x, o[1] = a.y, o[2] = a.z, o[3] = b.x, o[4] = b.y, o[5] = b.z, o[6] = a.w;
    }
  }
}
int photon_compact_blocks(long long total) { return (int)((total + kCompactBlock - 1) / kCompactBlock); }
void launch_photon_compact(const float4* out_a, const float4* out_b, long long total, int npaths, unsigned* block_count,
                           unsigned long long* light_count, unsigned* hist20, float* out7, long long capacity,
                           cudaStream_t st) {
  const int nb = photon_compact_blocks(total);
  if (nb < 1) return;
  k_photon_count<<<nb, kCompactBlock, 0, st>>>(out_b, total, npaths, block_count, light_count, hist20);
  k_scan_u32<<<1, 1024, 0, st>>>(block_count, nb);
  k_photon_compact<<<nb, kCompactBlock, 0, st>>>(out_a, out_b, total, block_count, out7, capacity);
}

// Splice the all-gathered shards into the single-process order: for every light, rank 0's particles, then rank 1's, ...
// seg_src[i] / seg_dst[i]: first particle of segment i = (light, rank) in the gathered buffer / in the output
// (seg_dst has one more entry: the total); segments are few (lights x ranks), found by a linear scan.
__global__ void k_photon_splice(const float* __restrict__ gathered, const long long* __restrict__ seg_src,
                                const long long* __restrict__ seg_dst, int nseg, float* out7) {
  const long long total = seg_dst[nseg];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 7 * total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / 7;
    const int c = (int)(i - 7 * p);
    int s = 0;
    while (s + 1 < nseg && seg_dst[s + 1] <= p) s++;
    out7[i] = gathered[7 * (seg_src[s] + (p - seg_dst[s])) + c];
  }
}
void launch_photon_splice(const float* gathered, const long long* seg_src, const long long* seg_dst, int nseg,
                          long long total, float* out7, cudaStream_t st) {
  if (total < 1) return;
  long long blocks = (7 * total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_photon_splice<<<(int)blocks, 256, 0, st>>>(gathered, seg_src, seg_dst, nseg, out7);
}
// particles (7 floats each) in kd order -> the two float4 arrays the gather reads
__global__ void k_photon_unpack(const float* __restrict__ p7, long long n, float4* kd_pos, float4* kd_dir) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* a = p7 + 7 * i;
  kd_pos[i] = make_float4(a[0], a[1], a[2], a[6]);
  kd_dir[i] = make_float4(a[3], a[4], a[5], 0.f);
}
void launch_photon_unpack(const float* p7, long long n, float4* kd_pos, float4* kd_dir, cudaStream_t st) {
  if (n < 1) return;
  k_photon_unpack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p7, n, kd_pos, kd_dir);
}

}  // namespace rtb
