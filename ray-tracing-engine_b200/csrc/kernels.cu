// kernels.cu -- the sm_100a kernels of the render hot path.
//
//   k_segment<MODE,PHOTON>  one path segment of the wavefront: (seg 0: jitter + camera ray) ->
//                           nearest-hit BVH traversal -> shade (3 any-hit shadow rays + microfacet BSDF,
//                           or k-nearest-photon gather) -> bounce sample -> compacted ray queue
//                           replaces Renderer.cpp:106-201,33-104 / RayTracer.h:27-53,95-117
//   k_resolve / k_scatter   ordered per-pixel accumulation of the clamped samples (Renderer.cpp:254-258)
//   k_emit                  photon emission + Russian-roulette random walk (PhotonMap.h:14-50,92-155)
//   k_trace / k_knn / k_bsdf parity hooks over caller-supplied batches
//
// No tensor-core work exists on this path (nothing is a dense contraction): the kernels are
// pointer-chasing traversals bounded by L1/L2 latency and the fp32 issue rate.
#include "kernels.h"

namespace rtb {

// ----------------------------------------------------------------------------------------------
// shading helpers shared by the segment and emission kernels
// ----------------------------------------------------------------------------------------------
template <bool ANY>
RT_DI bool trace(const DScene& S, float3 o, float3 d, int* stack, int brute, HitRec& h) {
  if (brute) return brute_trace<ANY>(S, o, d, h);
  return bvh_traverse<ANY>(S, o, d, stack, kBlock, h);
}

// Renderer.cpp:33-61
RT_DI float3 shade_direct(const DScene& S, float3 dir, float3 n, float3 P, const DMaterial& m, Rng& g, int* stack,
                          int brute, unsigned& n_shadow) {
  float3 color = f3(0.f, 0.f, 0.f);
  const float3 wo = v_neg(dir);
  for (int l = 0; l < S.num_lights; l++) {
    const DLight& L = S.lights[l];
    float3 to_light = v_sub(light_rand_area_position(L, g), P);
    HitRec hs;
    n_shadow++;
    if (trace<true>(S, P, to_light, stack, brute, hs)) continue;  // any hit on (0,+inf), Renderer.cpp:52-55
    float3 bsdf = evaluate_color_response(m, n, to_light, wo);
    float3 radiance = light_evaluate(L, P);
    color = v_add(color, v_mul(radiance, bsdf));
  }
  return color;
}

// Renderer.cpp:63-104
RT_DI float3 shade_photon(const DScene& S, float3 dir, float3 n, float3 P, const DMaterial& m, int k, int num_photons,
                          unsigned long long& visits) {
  float hd[kMaxK];
  int hi[kMaxK];
  int kst[3 * kKdStack];
  KdHeap H{hd, hi, 1};
  kd_knearest(S, P, k, H, kst, 1, visits);
  float r = H.d(k - 1);  // farthest of the k (result is sorted ascending)
  float area = (float)__dmul_rn(__dmul_rn(3.141592653589793, (double)r), (double)r);
  float3 avg = f3(0.f, 0.f, 0.f);
  float cnt = 0.f;
  for (int j = 0; j < k; j++) {
    avg = v_add(avg, f3(__ldg(S.kd_dir + H.i(j))));
    cnt = __fadd_rn(cnt, 1.f);
  }
  float rad = __fmul_rn(__fdiv_rn(__fdiv_rn(cnt, area), (float)num_photons), 100.f);
  float3 bsdf = evaluate_color_response(m, n, v_norm(avg), v_neg(dir));
  return v_scl(bsdf, rad);
}

// ----------------------------------------------------------------------------------------------
// k_segment
// ----------------------------------------------------------------------------------------------
template <int MODE, bool PHOTON>
__global__ void __launch_bounds__(kBlock) k_segment(const RenderArgs A, const int seg) {
  __shared__ int s_stack[kStackDepth * kBlock];
  int* stack = s_stack + threadIdx.x;
  const DScene& S = A.scene;
  const unsigned n = seg == 0 ? (unsigned)A.npix * (unsigned)A.nsamp : A.q_count[seg - 1];
  const float4* qo_in = A.q_o[(seg + 1) & 1];
  const float4* qd_in = A.q_d[(seg + 1) & 1];
  float4* qo_out = A.q_o[seg & 1];
  float4* qd_out = A.q_d[seg & 1];
  const unsigned lane = threadIdx.x & 31u;
  unsigned n_near = 0, n_shadow = 0, n_knn = 0;
  unsigned long long n_visits = 0;

  for (unsigned base = blockIdx.x * kBlock; base < n; base += gridDim.x * kBlock) {
    const unsigned i = base + threadIdx.x;
    bool push = false;
    float3 next_o = f3(0, 0, 0), next_d = f3(0, 0, 0);
    unsigned p = 0;
    if (i < n) {
      float3 o, d;
      if (seg == 0) {
        p = i;
      } else {
        float4 a = qo_in[i], b = qd_in[i];
        o = f3(a);
        d = f3(b);
        p = (unsigned)__float_as_int(a.w);
      }
      const int sl = p / (unsigned)A.npix;
      const int pl = p - sl * A.npix;
      const int pixel = __ldg(A.pix_map + pl);
      const int sample = A.s0 + sl;
      Rng g;
      const uint64_t key =
          stream_key(A.seed_mixed, kDomainPixel, (uint64_t)sample * ((uint64_t)A.width * A.height) + (uint64_t)pixel);
      if (seg == 0) {
        g.init(key, 0);
        float sx, sy;
        jitter_sample(g, sample, A.jitter_d, sx, sy);  // Renderer.cpp:229
        const int y = pixel / A.width, x = pixel - y * A.width;
        camera_ray(S.cam, x, y, sx, sy, A.width, A.height, o, d);  // Renderer.cpp:233
      } else {
        g.init(key, 4u + (PHOTON ? 4u : 10u) * (unsigned)seg);
      }
      HitRec h;
      n_near++;
      const bool found = trace<false>(S, o, d, stack, A.brute, h);  // h.t > 0 by construction
      float3 c = f3(0.f, 0.f, 0.f);
      if (found) {
        float3 nrm, P;
        int mesh;
        hit_geometry(S, h, nrm, P, mesh);
        const DMaterial m = S.mats[mesh];
        if (PHOTON) {
          n_knn++;
          c = shade_photon(S, d, nrm, P, m, A.k, A.num_photons, n_visits);
        } else {
          c = shade_direct(S, d, nrm, P, m, g, stack, A.brute, n_shadow);
        }
        if (MODE == 1 && seg < 2) {
          next_d = hsphere_uniform_sample(g, nrm);  // Renderer.cpp:164-166
          next_o = P;
          push = true;
        }
      }
      // Renderer.cpp:143-170 unrolled: colour = c0 + (c1 + c2); a miss at depth d ends the sum there.
      if (seg == 0) {
        float3 out = (MODE == 0 || !found) ? normalize_color(c) : c;
        A.col0[p] = make_float4(out.x, out.y, out.z, found ? 1.f : 0.f);
      } else if (seg == 1) {
        if (found) {
          A.col1[p] = make_float4(c.x, c.y, c.z, 0.f);
        } else {
          float4 c0 = A.col0[p];
          float3 out = normalize_color(f3(c0));
          A.col0[p] = make_float4(out.x, out.y, out.z, c0.w);
        }
      } else {
        float4 c0 = A.col0[p], c1 = A.col1[p];
        float3 out = normalize_color(v_add(f3(c0), v_add(f3(c1), c)));
        A.col0[p] = make_float4(out.x, out.y, out.z, c0.w);
      }
    }
    if (MODE == 1 && seg < 2) {  // warp-aggregated push into the next segment's queue
      const unsigned mask = __ballot_sync(0xffffffffu, push);
      if (mask) {
        unsigned slot0 = 0;
        if (lane == (unsigned)(__ffs(mask) - 1)) slot0 = atomicAdd(A.q_count + seg, (unsigned)__popc(mask));
        slot0 = __shfl_sync(0xffffffffu, slot0, __ffs(mask) - 1);
        if (push) {
          const unsigned slot = slot0 + __popc(mask & ((1u << lane) - 1u));
          qo_out[slot] = make_float4(next_o.x, next_o.y, next_o.z, __int_as_float((int)p));
          qd_out[slot] = make_float4(next_d.x, next_d.y, next_d.z, 0.f);
        }
      }
    }
  }
  // counters: one atomic per warp per counter
  for (int off = 16; off > 0; off >>= 1) {
    n_near += __shfl_xor_sync(0xffffffffu, n_near, off);
    n_shadow += __shfl_xor_sync(0xffffffffu, n_shadow, off);
    n_knn += __shfl_xor_sync(0xffffffffu, n_knn, off);
    n_visits += __shfl_xor_sync(0xffffffffu, n_visits, off);
  }
  if (lane == 0) {
    if (n_near) atomicAdd(A.counters + kCntNearest, (unsigned long long)n_near);
    if (n_shadow) atomicAdd(A.counters + kCntShadow, (unsigned long long)n_shadow);
    if (n_knn) atomicAdd(A.counters + kCntKnn, (unsigned long long)n_knn);
    if (n_visits) atomicAdd(A.counters + kCntKdVisits, n_visits);
  }
}

void launch_segment(const RenderArgs& a, int seg, int grid_ctas, cudaStream_t st) {
  if (a.mode == 0) {
    if (a.photon)
      k_segment<0, true><<<grid_ctas, kBlock, 0, st>>>(a, seg);
    else
      k_segment<0, false><<<grid_ctas, kBlock, 0, st>>>(a, seg);
  } else {
    if (a.photon)
      k_segment<1, true><<<grid_ctas, kBlock, 0, st>>>(a, seg);
    else
      k_segment<1, false><<<grid_ctas, kBlock, 0, st>>>(a, seg);
  }
}

int segment_ctas_per_sm(int mode, int photon) {
  int n = 0;
  if (mode == 0)
    photon ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_segment<0, true>, kBlock, 0)
           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_segment<0, false>, kBlock, 0);
  else
    photon ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_segment<1, true>, kBlock, 0)
           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_segment<1, false>, kBlock, 0);
  return n < 1 ? 1 : n;
}

// ----------------------------------------------------------------------------------------------
// ordered accumulation: updateImage(x,y) += colorResponse for samples in index order
// ----------------------------------------------------------------------------------------------
__global__ void k_resolve(const float4* __restrict__ col0, int npix, int nsamp, float4* acc_rgb, int* acc_cnt) {
  int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= npix) return;
  float4 acc = acc_rgb[pl];
  int cnt = acc_cnt[pl];
  for (int s = 0; s < nsamp; s++) {
    float4 c = col0[(size_t)s * npix + pl];
    acc.x = __fadd_rn(acc.x, c.x);
    acc.y = __fadd_rn(acc.y, c.y);
    acc.z = __fadd_rn(acc.z, c.z);
    cnt += (c.w != 0.f) ? 1 : 0;  // Renderer.cpp:255-257
  }
  acc_rgb[pl] = acc;
  acc_cnt[pl] = cnt;
}
void launch_resolve(const float4* col0, int npix, int nsamp, float4* acc_rgb, int* acc_cnt, cudaStream_t st) {
  k_resolve<<<(npix + 255) / 256, 256, 0, st>>>(col0, npix, nsamp, acc_rgb, acc_cnt);
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

// ----------------------------------------------------------------------------------------------
// parity hooks
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_trace(const DScene S, const float* __restrict__ rays6, long long n,
                                                  int* tri, float* uvt, int brute, int any, unsigned char* occluded) {
  __shared__ int s_stack[kStackDepth * kBlock];
  int* stack = s_stack + threadIdx.x;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    float3 o = f3(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]);
    float3 d = f3(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]);
    HitRec h;
    if (any) {
      occluded[i] = trace<true>(S, o, d, stack, brute, h) ? 1 : 0;
    } else {
      bool f = trace<false>(S, o, d, stack, brute, h);
      tri[i] = f ? h.gid : -1;
      uvt[3 * i] = f ? h.u : 0.f;
      uvt[3 * i + 1] = f ? h.v : 0.f;
      uvt[3 * i + 2] = f ? h.t : 0.f;
    }
  }
}
void launch_trace_rays(const DScene& s, const float* rays6, long long n, int* tri, float* uvt, int brute, int any,
                       unsigned char* occluded, cudaStream_t st) {
  long long blocks = (n + kBlock - 1) / kBlock;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_trace<<<(int)blocks, kBlock, 0, st>>>(s, rays6, n, tri, uvt, brute, any, occluded);
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

__global__ void __launch_bounds__(kBlock) k_knn(const DScene S, const float* __restrict__ q3, long long n, int k,
                                                int* node_index, unsigned long long* counters) {
  long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  float hd[kMaxK];
  int hi[kMaxK];
  int kst[3 * kKdStack];
  KdHeap H{hd, hi, 1};
  unsigned long long visits = 0;
  kd_knearest(S, f3(q3[3 * i], q3[3 * i + 1], q3[3 * i + 2]), k, H, kst, 1, visits);
  for (int j = 0; j < k; j++) node_index[i * k + j] = H.i(j);
  atomicAdd(counters + kCntKdVisits, visits);
  atomicAdd(counters + kCntKnn, 1ull);
}
void launch_knn(const DScene& s, const float* q3, long long n, int k, int* node_index, unsigned long long* counters,
                cudaStream_t st) {
  k_knn<<<(int)((n + kBlock - 1) / kBlock), kBlock, 0, st>>>(s, q3, n, k, node_index, counters);
}

// ----------------------------------------------------------------------------------------------
// k_emit: one thread per photon path (PhotonMap.h:19-44 emission, :92-155 random walk).
// out_a = (position, weight), out_b = (incomeDirection, status bits): bit0 stored,
// bits 8.. = 1 + depth of the Russian-roulette kill (0: not counted in the histogram).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_emit(const DScene S, uint64_t seed_mixed, int per_light, float light_pdf,
                                                 int first_path, int npaths, int brute, float4* out_a, float4* out_b,
                                                 unsigned long long* counters) {
  __shared__ int s_stack[kStackDepth * kBlock];
  int* stack = s_stack + threadIdx.x;
  const long long total = (long long)S.num_lights * npaths;
  unsigned n_rays = 0;
  for (long long q = (long long)blockIdx.x * kBlock + threadIdx.x; q < total; q += (long long)gridDim.x * kBlock) {
    const int li = (int)(q / npaths);
    const int path = first_path + (int)(q - (long long)li * npaths);
    const DLight& L = S.lights[li];
    Rng g;
    g.init(stream_key(seed_mixed, kDomainPhoton, (uint64_t)li * (uint64_t)per_light + (uint64_t)path), 0);
    float3 o = light_rand_area_position(L, g);
    float3 d = hsphere_uniform_sample(g, L.normal);
    float pdf0 = v_dot(v_norm(d), v_norm(L.normal));
    float weight = __fdiv_rn(light_radiance(L, o), __fmul_rn(pdf0, light_pdf));
    float3 ppos = f3(0.f, 0.f, 0.f), pdir = f3(0.f, 0.f, 0.f);
    bool exit = false, stored = false;
    int hist = 0;
    for (int depth = 0;; depth++) {
      if (exit) {  // PhotonMap.h:94-97
        stored = true;
        hist = depth;  // 1 + (depth-1)
        break;
      }
      if (depth >= 20) break;  // PhotonMap.h:98: dropped
      HitRec h;
      n_rays++;
      if (!trace<false>(S, o, d, stack, brute, h)) {  // PhotonMap.h:109-112
        stored = depth != 0;
        break;
      }
      float3 nrm, P;
      int mesh;
      hit_geometry(S, h, nrm, P, mesh);
      const DMaterial m = S.mats[mesh];
      ppos = P;
      pdir = v_neg(d);
      float3 rd = hsphere_uniform_sample(g, nrm);
      float3 refl = v_sub(d, v_scl(nrm, __fmul_rn(2.f, v_dot(d, nrm))));
      float bsdf = v_len(evaluate_color_response(m, nrm, d, rd));
      float pdf = __fdiv_rn(__fadd_rn(v_dot(v_norm(rd), v_norm(refl)), 1.f), 2.f);
      weight = __fmul_rn(weight, __fdiv_rn(bsdf, pdf));
      float cont = fminf(weight, 1.f);
      if (g.uniform_f(0.f, 1.f) > cont)  // PhotonMap.h:144-150
        exit = true;
      else
        weight = __fdiv_rn(weight, cont);
      o = P;
      d = rd;
    }
    out_a[q] = make_float4(ppos.x, ppos.y, ppos.z, weight);
    out_b[q] = make_float4(pdir.x, pdir.y, pdir.z, __int_as_float((stored ? 1 : 0) | (hist << 8)));
  }
  for (int off = 16; off > 0; off >>= 1) n_rays += __shfl_xor_sync(0xffffffffu, n_rays, off);
  if ((threadIdx.x & 31) == 0 && n_rays) atomicAdd(counters + kCntPhotonRays, (unsigned long long)n_rays);
}
void launch_emit(const DScene& s, uint64_t seed_mixed, int per_light, float light_pdf, int first_path, int npaths,
                 int brute, float4* out_a, float4* out_b, unsigned long long* counters, cudaStream_t st) {
  long long total = (long long)s.num_lights * npaths;
  long long blocks = (total + kBlock - 1) / kBlock;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_emit<<<(int)blocks, kBlock, 0, st>>>(s, seed_mixed, per_light, light_pdf, first_path, npaths, brute, out_a, out_b,
                                         counters);
}

}  // namespace rtb
