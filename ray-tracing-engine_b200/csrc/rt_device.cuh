// rt_device.cuh -- device-side building blocks of the render hot path (sm_100a).
//
// Everything that decides WHICH triangle a ray hits, WHERE the hit point is and WHETHER a shadow ray
// is blocked is strict IEEE binary32 in the reference's operation order with no fused multiply-add
// (occlusion decisions at t ~ 1e-7 are bit-chaotic).  Those routines
// use the round-to-nearest intrinsics (__fmul_rn/__fadd_rn/...), which nvcc never contracts, so the
// result does not depend on -fmad; the file is compiled with -fmad=false as a second line of defence.
// Reference citations are file:line under /root/reference/source.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace rtb {

// ------------------------------------------------------------------------------------------------
// strict float3 algebra (Vec3.h)
// ------------------------------------------------------------------------------------------------
#define RT_DI __device__ __forceinline__

RT_DI float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
RT_DI float3 f3(float4 a) { return make_float3(a.x, a.y, a.z); }
RT_DI float3 v_add(float3 a, float3 b) { return f3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
RT_DI float3 v_sub(float3 a, float3 b) { return f3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
RT_DI float3 v_neg(float3 a) { return f3(-a.x, -a.y, -a.z); }
RT_DI float3 v_mul(float3 a, float3 b) { return f3(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z)); }
RT_DI float3 v_scl(float3 a, float s) { return f3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
RT_DI float3 v_dvs(float3 a, float s) { return f3(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)); }
// Vec3.h:220-223  (a0*b0 + a1*b1) + a2*b2
RT_DI float v_dot(float3 a, float3 b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
// Vec3.h:225-232
RT_DI float3 v_cross(float3 a, float3 b) {
  return f3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
            __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// Vec3.h:165-167: (float)sqrt((double)x).  A correctly rounded binary32 sqrt gives the same value
// (double rounding is innocuous for sqrt when the wide format has >= 2*24+2 bits).
RT_DI float v_len(float3 a) { return __fsqrt_rn(v_dot(a, a)); }
RT_DI float v_dist(float3 a, float3 b) { return v_len(v_sub(a, b)); }
// Vec3.h:170-178: zero stays zero; otherwise multiply by 1/len (not divide by len).
RT_DI float3 v_norm(float3 a) {
  float l = v_len(a);
  if (l == 0.0f) return a;
  float inv = __frcp_rn(l);  // == 1.0f / l, correctly rounded
  return v_scl(a, inv);
}
// Vec3.h:180-199
RT_DI void v_two_orthogonals(float3 n, float3& u, float3& v) {
  if (fabsf(n.x) < fabsf(n.y)) {
    if (fabsf(n.x) < fabsf(n.z))
      u = f3(0.f, -n.z, n.y);
    else
      u = f3(-n.y, n.x, 0.f);
  } else {
    if (fabsf(n.y) < fabsf(n.z))
      u = f3(n.z, 0.f, -n.x);
    else
      u = f3(-n.y, n.x, 0.f);
  }
  v = v_cross(n, u);
}

// ------------------------------------------------------------------------------------------------
// counter-based random stream (DESIGN.md "RNG contract"); word -> uniform arithmetic of libstdc++'s
// generate_canonical for a 32-bit engine
// ------------------------------------------------------------------------------------------------
constexpr uint64_t kGolden = 0x9E3779B97F4A7C15ull;
constexpr uint64_t kDomainPixel = 1ull;
constexpr uint64_t kDomainPhoton = 2ull;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}
__host__ __device__ __forceinline__ uint64_t stream_key(uint64_t seed_mixed, uint64_t domain, uint64_t index) {
  return mix64(seed_mixed ^ ((domain << 56) | index));  // seed_mixed = mix64(seed + kGolden), host-computed
}

struct Rng {
  uint64_t key;
  uint64_t pair;      // cached pair(ctr/2)
  uint32_t ctr;       // next word index
  uint32_t pair_idx;  // which pair is cached (0xffffffff = none)
  RT_DI void init(uint64_t k, uint32_t first_word) {
    key = k;
    ctr = first_word;
    pair_idx = 0xffffffffu;
    pair = 0;
  }
  RT_DI uint32_t word() {
    uint32_t j = ctr >> 1;
    if (j != pair_idx) {
      pair = mix64(key + (uint64_t)(j + 1u) * kGolden);
      pair_idx = j;
    }
    uint32_t w = (ctr & 1u) ? (uint32_t)(pair >> 32) : (uint32_t)pair;
    ctr++;
    return w;
  }
  RT_DI float canonical_f() {
    float f = __fmul_rn(__uint2float_rn(word()), 2.3283064365386963e-10f);  // float(w) / 2^32 (exact scaling)
    if (f >= 1.0f) f = 0.99999994f;                                        // nextafterf(1, 0)
    return f;
  }
  RT_DI double canonical_d() {
    double lo = (double)word();
    double hi = (double)word();
    double g = __dmul_rn(__dadd_rn(lo, __dmul_rn(hi, 4294967296.0)), 5.421010862427522e-20);  // / 2^64
    if (g >= 1.0) g = 0.99999999999999989;  // nextafter(1.0, 0.0)
    return g;
  }
  // uniform_real_distribution: canonical * (b - a) + a
  RT_DI float uniform_f(float a, float b) { return __fadd_rn(__fmul_rn(canonical_f(), __fsub_rn(b, a)), a); }
  RT_DI double uniform_d0(double b) { return __dadd_rn(__dmul_rn(canonical_d(), b), 0.0); }  // a == 0.0
};

// ------------------------------------------------------------------------------------------------
// device scene
// ------------------------------------------------------------------------------------------------
struct DMaterial {  // Material.h:62-64 + per-material constants of Material.h:25-69, precomputed by
  float kd, alpha;    // make_material() with exactly the reference's operations (so they are bit-identical
  float3 albedo, f0;  // to recomputing them per call, as the reference does)
  float3 diffuse_kd;   // m_kd * (m_albedo / float(M_PI))          Material.h:27,38
  float one_minus_kd;  // 1 - m_kd                                 Material.h:28
  float a2, a2m1;      // alpha*alpha, alpha*alpha - 1             Material.h:47-48
  float k, one_minus_k;  // float(alpha*sqrt(2/pi)), 1 - k         Material.h:67-68
  float3 one_minus_f0;   // (1,1,1) - F0                           Material.h:52
};
__host__ __device__ inline DMaterial make_material(float kd, float alpha, float3 albedo, float3 f0) {
  DMaterial m;
  m.kd = kd;
  m.alpha = alpha;
  m.albedo = albedo;
  m.f0 = f0;
#ifdef __CUDA_ARCH__
  const float pi_f = 3.14159274f;
  m.diffuse_kd = make_float3(__fmul_rn(__fdiv_rn(albedo.x, pi_f), kd), __fmul_rn(__fdiv_rn(albedo.y, pi_f), kd),
                             __fmul_rn(__fdiv_rn(albedo.z, pi_f), kd));
  m.one_minus_kd = __fsub_rn(1.f, kd);
  m.a2 = __fmul_rn(alpha, alpha);
  m.a2m1 = __fsub_rn(m.a2, 1.f);
  m.k = (float)__dmul_rn((double)alpha, 0.7978845608028654);
  m.one_minus_k = __fsub_rn(1.f, m.k);
  m.one_minus_f0 = make_float3(__fsub_rn(1.f, f0.x), __fsub_rn(1.f, f0.y), __fsub_rn(1.f, f0.z));
#else
  // host: x86-64 SSE2 binary32/binary64, compiled with -ffp-contract=off (volatile blocks re-association)
  volatile float pi_f = 3.14159274f;
  volatile float dx = albedo.x / pi_f, dy = albedo.y / pi_f, dz = albedo.z / pi_f;
  m.diffuse_kd = make_float3(dx * kd, dy * kd, dz * kd);
  m.one_minus_kd = 1.f - kd;
  volatile float a2 = alpha * alpha;
  m.a2 = a2;
  m.a2m1 = a2 - 1.f;
  volatile double kk = (double)alpha * 0.7978845608028654;
  m.k = (float)kk;
  m.one_minus_k = 1.f - m.k;
  m.one_minus_f0 = make_float3(1.f - f0.x, 1.f - f0.y, 1.f - f0.z);
#endif
  return m;
}
struct DLight {  // LightSource.h:61-65
  float3 position, color, normal, vertical, horizontal;
  float intensity, side, ac, al, aq, factor;
};
struct DCamera {  // Camera.h:34-40
  float3 position, lower_left, horizontal, vertical;
};

constexpr int kMaxLights = 8;     // lights kept in kernel-parameter space; further ones are read from lights_ext
constexpr int kMaxRoots = 16;    // scenes with more meshes than this enter through the top-level tree
constexpr int kStackDepth = 40;  // BVH traversal stack entries per ray (host checks bvh depth <= this)
constexpr int kBlock = 128;      // threads per CTA of the wavefront kernels
constexpr int kMaxK = 64;        // largest k of the ascending-array k-NN (its tie fallback keeps k candidates in local memory)
constexpr int kKdStack = 32;     // kd-tree recursion depth bound (host checks)

struct DScene {
  const float4* __restrict__ nodes;     // 4 x float4 per BVH node (see bvh_build.h)
  const float4* __restrict__ tris;      // 3 x float4 per triangle slot, BVH leaf order: (p0,gid) (e1,-) (e2,-)
  const int4* __restrict__ tri_vidx;    // per GLOBAL triangle: v0, v1, v2 (global vertex ids), mesh
  const float4* __restrict__ pos;       // per vertex
  const float4* __restrict__ nrm;       // per vertex
  const DMaterial* __restrict__ mats;   // per mesh
  DLight lights[kMaxLights];
  const DLight* __restrict__ lights_ext;  // lights kMaxLights, kMaxLights+1, ... (device memory); null when <= kMaxLights
  int num_lights;
  int num_tris;
  DCamera cam;
  // per-mesh BVH roots (mesh order): lo.xyz, hi.xyz, hi.w = node/leaf reference; 0 roots = start at node 0
  int num_roots;
  float4 root_lo[kMaxRoots], root_hi[kMaxRoots];
  // photon map: kd-tree in the array order of kdtree::make_tree (kdtree.h:60-69), links implicit
  const float4* __restrict__ kd_pos;  // xyz = position, w = weight
  const float4* __restrict__ kd_dir;  // xyz = incomeDirection
  int kd_count;
};

struct HitRec {
  float t, u, v;
  int gid;  // global triangle index, scene order
};
// light l of the scene (Scene::lightsources()[l], any count: Renderer.cpp:49, PhotonMap.h:24)
RT_DI const DLight& light_at(const DScene& S, int l) { return l < kMaxLights ? S.lights[l] : S.lights_ext[l - kMaxLights]; }

// ------------------------------------------------------------------------------------------------
// Ray.cpp:9-24  Moller-Trumbore with the absolute epsilon on det.  e1/e2 are the host-precomputed
// p1-p0 / p2-p0 (the same binary32 subtractions the reference performs per test).  Early exits skip
// work the reference would discard; accepted hits get bit-identical u, v, t.
// ------------------------------------------------------------------------------------------------
RT_DI bool mt_intersect(float3 o, float3 d, float3 p0, float3 e1, float3 e2, float& u, float& v, float& t) {
  float3 pvec = v_cross(d, e2);
  float det = v_dot(e1, pvec);
  if (fabsf(det) < 0.000001f) return false;
  float inv_det = __frcp_rn(det);  // == 1.0f / det
  float3 tvec = v_sub(o, p0);
  u = __fmul_rn(v_dot(tvec, pvec), inv_det);
  if (u < 0.f || u > 1.f) return false;
  float3 qvec = v_cross(tvec, e1);
  v = __fmul_rn(v_dot(d, qvec), inv_det);
  if (!(v >= 0.f && __fadd_rn(u, v) <= 1.f)) return false;
  t = __fmul_rn(v_dot(e2, qvec), inv_det);
  return true;
}

// RayTracer.h:40 acceptance, made order-independent: the reference scans in scene order with a
// strict `dt < closest`, i.e. it returns the minimum t and, among exactly equal t, the lowest index.
RT_DI void accept_nearest(HitRec& h, float t, float u, float v, int gid) {
  if (t > 0.f && t < FLT_MAX && (t < h.t || (t == h.t && gid < h.gid))) {
    h.t = t;
    h.u = u;
    h.v = v;
    h.gid = gid;
  }
}

// Conservative slab test against a padded box (see host_build.h for the padding argument).  NaNs from
// 0*inf drop out of fminf/fmaxf.  A child is visited when its interval overlaps [0, best_t].
RT_DI bool box_hit(float3 lo, float3 hi, float3 o, float3 inv, float best_t, float& tnear) {
  float t0x = (lo.x - o.x) * inv.x, t1x = (hi.x - o.x) * inv.x;
  float t0y = (lo.y - o.y) * inv.y, t1y = (hi.y - o.y) * inv.y;
  float t0z = (lo.z - o.z) * inv.z, t1z = (hi.z - o.z) * inv.z;
  float tmin = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
  float tmax = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
  tnear = tmin;
  return (tmin <= tmax) && (tmax >= 0.f) && (tmin <= best_t);
}

// The same test with one fused multiply-add per plane: t = lo*inv - o*inv (oinv = o*inv precomputed per
// ray).  The traversal only has to be conservative, not bit-exact, and the padding covers the different
// rounding.  inv MUST come from safe_inv(): with inv = +-inf the form lo*inf - o*inf gives inf - inf = NaN
// on ONE plane of a slab the origin is inside of, and fmin/fmax then resolve it as a miss (found by the
// N=128 parity test: 3 of 22.6 M primary rays have an exactly-zero direction component).
// The reciprocal itself only steers the (conservative) slab test, so the 1-ulp hardware approximation is enough: it
// adds 2^-23 relative to a t whose error budget is pad / |t d| ~ 2^-14 (host_build.h), and saves the ~10-instruction
// IEEE division sequence three times per ray refill.  d = +-0 or denormal gives +-inf, clamped like before.
RT_DI float safe_inv(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return fminf(fmaxf(r, -1e30f), 1e30f);
}
// The three conditions (tmin <= tmax, tmax >= 0, tmin <= best_t) are one comparison of folded bounds,
// max(tmin, 0) <= min(tmax, best_t)  (best_t >= 0 always): two FMNMX and one FSETP instead of three FSETP and the
// predicate logic -- 2-3 of the ~19 instructions per box in an issue-bound loop.  UNBOUNDED (any-hit: best_t stays
// +inf until the ray ends) drops the min with best_t as well.
template <bool UNBOUNDED = false>
RT_DI bool box_hit_fma(float3 lo, float3 hi, float3 inv, float3 oinv, float best_t, float& tnear) {
  float t0x = __fmaf_rn(lo.x, inv.x, -oinv.x), t1x = __fmaf_rn(hi.x, inv.x, -oinv.x);
  float t0y = __fmaf_rn(lo.y, inv.y, -oinv.y), t1y = __fmaf_rn(hi.y, inv.y, -oinv.y);
  float t0z = __fmaf_rn(lo.z, inv.z, -oinv.z), t1z = __fmaf_rn(hi.z, inv.z, -oinv.z);
  float tmin = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
  float tmax = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
  tnear = tmin;
  const float enter = fmaxf(tmin, 0.f), leave = UNBOUNDED ? tmax : fminf(tmax, best_t);
  return enter <= leave;
}

// Nearest-hit (ANY=false) or any-hit (ANY=true) traversal of the flattened BVH.
// `stack` points at this thread's column of a [kStackDepth][blockDim.x] shared-memory array.
template <bool ANY>
RT_DI bool bvh_traverse(const DScene& S, float3 o, float3 d, int* stack, int stride, HitRec& h) {
  h.t = FLT_MAX;
  h.u = 0.f;
  h.v = 0.f;
  h.gid = 0x7fffffff;
  const float3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  int sp = 0;
  int cur = 0;
  for (;;) {
    const float4 n0 = __ldg(S.nodes + 4 * cur);
    const float4 n1 = __ldg(S.nodes + 4 * cur + 1);
    const float4 n2 = __ldg(S.nodes + 4 * cur + 2);
    const float4 n3 = __ldg(S.nodes + 4 * cur + 3);
    float tn0, tn1;
    bool h0 = box_hit(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), o, inv, h.t, tn0);
    bool h1 = box_hit(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), o, inv, h.t, tn1);
    int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
#pragma unroll
    for (int side = 0; side < 2; side++) {
      bool& hs = side ? h1 : h0;
      int c = side ? c1 : c0;
      if (hs && c < 0) {  // leaf child: one triangle
        int slot = ~c;
        float4 A = __ldg(S.tris + 3 * slot), B = __ldg(S.tris + 3 * slot + 1), C = __ldg(S.tris + 3 * slot + 2);
        float u, v, t;
        if (mt_intersect(o, d, f3(A), f3(B), f3(C), u, v, t)) {
          if (ANY) {
            if (t > 0.f && t < FLT_MAX) return true;
          } else {
            accept_nearest(h, t, u, v, __float_as_int(A.w));
          }
        }
        hs = false;
      }
    }
    if (h0 && h1) {
      bool first0 = tn0 <= tn1;
      stack[sp * stride] = first0 ? c1 : c0;
      sp++;
      cur = first0 ? c0 : c1;
    } else if (h0) {
      cur = c0;
    } else if (h1) {
      cur = c1;
    } else {
      if (sp == 0) break;
      sp--;
      cur = stack[sp * stride];
    }
  }
  return ANY ? false : (h.gid != 0x7fffffff);
}

// RayTracer.h:27-53 as written: scan every triangle (parity hook / tiny scenes).
template <bool ANY>
RT_DI bool brute_trace(const DScene& S, float3 o, float3 d, HitRec& h) {
  h.t = FLT_MAX;
  h.u = 0.f;
  h.v = 0.f;
  h.gid = 0x7fffffff;
  for (int s = 0; s < S.num_tris; s++) {
    float4 A = __ldg(S.tris + 3 * s), B = __ldg(S.tris + 3 * s + 1), C = __ldg(S.tris + 3 * s + 2);
    float u, v, t;
    if (mt_intersect(o, d, f3(A), f3(B), f3(C), u, v, t)) {
      if (ANY) {
        if (t > 0.f && t < FLT_MAX) return true;
      } else {
        accept_nearest(h, t, u, v, __float_as_int(A.w));
      }
    }
  }
  return ANY ? false : (h.gid != 0x7fffffff);
}

// ------------------------------------------------------------------------------------------------
// sampling
// ------------------------------------------------------------------------------------------------
// RayTracer.h:109-117 (sums formed in double, rounded once); jd = int(sqrt(float(N))) from the host
RT_DI void jitter_sample(Rng& g, int sample, int jd, float& x, float& y) {
  int j2 = sample / jd, i2 = sample % jd;
  double fd = (double)(float)jd;
  x = (float)__ddiv_rn(__dadd_rn((double)(float)i2, g.uniform_d0(1.0)), fd);
  y = (float)__ddiv_rn(__dadd_rn((double)(float)j2, g.uniform_d0(1.0)), fd);
}
// Camera.h:27-30 through Renderer.cpp:233-234
RT_DI void camera_ray(const DCamera& c, int x, int y, float sx, float sy, int w, int h, float3& o, float3& d) {
  float u = __fdiv_rn(__fadd_rn((float)x, sx), (float)w);
  float v = __fsub_rn(1.f, __fdiv_rn(__fadd_rn((float)y, sy), (float)h));
  o = c.position;
  d = v_norm(v_sub(v_add(v_add(c.lower_left, v_scl(c.horizontal, u)), v_scl(c.vertical, v)), c.position));
}
// ------------------------------------------------------------------------------------------------
// libm's binary32 sine and cosine, restated.  RayTracer.h:104-106 calls cos(phi), sin(phi), cos(theta), sin(theta)
// on floats, i.e. glibc's cosf/sinf -- which are NOT correctly rounded: since glibc 2.28 they are the ARM
// optimized-routines algorithm (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, s_sincosf.h, s_sincosf_data.c): reduce by
// pi/2 in binary64, a degree-7 / degree-8 polynomial in binary64, one rounding to binary32, 0.56 ulp.  A correctly
// rounded sine differs from that result for 0.36 % of the arguments in [0, 2 pi] and the cosine for 0.16 %; with four
// calls per bounce that alone made ~1-2 % of the -m 1 samples differ from the reference's in the last bit.  The
// algorithm is deterministic binary64 arithmetic, so it can be restated exactly: same table, same operation order,
// fused multiply-adds where the FMA build of libm (the variant x86-64 hosts with FMA select) has them.  Pinned on
// the host against this image's glibc 2.39 over every binary32 in [2^-15, 2 pi] (tests/test_host.py builds the same
// restatement in C) and on the device by the hsphere golden (tests/test_gpu_round2.py).
// Arguments here are in [0, 2 pi (1 + 3e-8)] or NaN (asin of a uniform that may exceed 1 by 3e-8, RayTracer.h:96).
// ------------------------------------------------------------------------------------------------
// the two polynomials of sinf_poly (s_sincosf.h): the sine one on (x, x^2), the cosine one on x^2 with the table's sign
RT_DI float libm_sin_poly(double x, double x2) {
  const double s1c = -0x1.555545995a603p-3, s2c = 0x1.1107605230bc4p-7, s3c = -0x1.994eb3774cf24p-13;
  const double x3 = __dmul_rn(x, x2);
  const double s1 = __fma_rn(x2, s3c, s2c);
  const double x7 = __dmul_rn(x3, x2);
  const double s = __fma_rn(x3, s1c, x);
  return (float)__fma_rn(x7, s1, s);
}
RT_DI float libm_cos_poly(double x2, bool second_table) {
  // __sincosf_table[0] and [1]: the cosine coefficients change sign in the second table, the sine ones do not
  const double sg = second_table ? -1.0 : 1.0;
  const double c0 = sg * 0x1p0, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5,
               c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
  const double x4 = __dmul_rn(x2, x2);
  const double q2 = __fma_rn(x2, c4, c3);
  const double q1 = __fma_rn(x2, c1, c0);
  const double x6 = __dmul_rn(x4, x2);
  const double c = __fma_rn(x4, c2, q1);
  return (float)__fma_rn(x6, q2, c);
}
// sinf(y) and cosf(y) of ONE argument: both calls reduce y the same way (same n, same remainder) and then evaluate
// one the sine polynomial and the other the cosine polynomial -- which one depends on the quadrant's parity.  Both
// polynomials are evaluated once, unconditionally (no divergence between lanes in different quadrants), and
// assigned by parity: bit for bit what two separate libm calls return.
RT_DI void libm_sincosf(float y, float& sin_y, float& cos_y) {
  const unsigned top = (__float_as_uint(y) >> 20) & 0x7ffu;  // abstop12
  const double x = (double)y;
  if (top < 0x3f4u) {  // |y| < pi/4
    if (top < 0x398u) {  // |y| < 2^-12
      sin_y = y;
      cos_y = 1.0f;
      return;
    }
    const double x2 = __dmul_rn(x, x);
    sin_y = libm_sin_poly(x, x2);
    cos_y = libm_cos_poly(x2, false);
    return;
  }
  if (top >= 0x42fu) {  // |y| >= 120, inf, NaN: never a finite argument on this path
    sin_y = sinf(y);
    cos_y = cosf(y);
    return;
  }
  // reduce_fast without TOINT_INTRINSICS: hpi_inv is 2/pi * 2^24
  const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
  const int n = (__double2int_rz(r) + 0x800000) >> 24;
  const double xr = __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
  const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // sign[] = {1, -1, -1, 1}
  const double xs = __dmul_rn(xr, sign), x2 = __dmul_rn(xr, xr);
  const float s = libm_sin_poly(xs, x2), c = libm_cos_poly(x2, (n & 2) != 0);
  sin_y = (n & 1) ? c : s;  // sinf: sinf_poly(x*s, x*x, p, n);  cosf: sinf_poly(x*s, x*x, p, n ^ 1)
  cos_y = (n & 1) ? s : c;
}

// RayTracer.h:95-107 with maxRayAngle = float(pi/2).  asin in binary64 like the reference (the result is rounded to
// binary32 at once, which hides the last-bit differences between libm's and CUDA's binary64 asin except with
// probability ~1e-8); cos/sin are libm's binary32 routines restated above.
RT_DI float3 hsphere_uniform_sample(Rng& g, float3 normal) {
  const double hi = 1.0000000278275352;  // 2*float(pi/2)/pi
  normal = v_norm(normal);
  float3 v1, v2;
  v_two_orthogonals(normal, v1, v2);
  v1 = v_norm(v1);
  v2 = v_norm(v2);
  float theta = (float)asin(g.uniform_d0(hi));
  float phi = (float)__dmul_rn(6.283185307179586, g.uniform_d0(hi));
  float sp, cp, st, ct;
  libm_sincosf(phi, sp, cp);
  libm_sincosf(theta, st, ct);
  float3 direction = v_norm(v_add(v_scl(v1, cp), v_scl(v2, sp)));
  return v_norm(v_add(v_scl(normal, ct), v_scl(direction, st)));
}
// LightSource.h:46-49: the first draw scales m_horizontal, the second m_vertical (g++ evaluates the
// right operand of the outer + first; pinned against the reference build by the oracle tests).
RT_DI float3 light_rand_area_position(const DLight& l, Rng& g) {
  float a = g.uniform_f(-l.side, l.side);
  float b = g.uniform_f(-l.side, l.side);
  return v_add(v_add(l.position, v_scl(l.vertical, b)), v_scl(l.horizontal, a));
}
// LightSource.h:51-54
RT_DI float light_radiance(const DLight& l, float3 p) {
  float d = v_dist(p, l.position);
  return __fdiv_rn(l.intensity, __fadd_rn(__fadd_rn(l.ac, __fmul_rn(l.al, d)), __fmul_rn(__fmul_rn(l.aq, d), d)));
}
// LightSource.h:56-59
RT_DI float3 light_evaluate(const DLight& l, float3 p) { return v_scl(v_scl(l.color, l.factor), light_radiance(l, p)); }

// ------------------------------------------------------------------------------------------------
// Material.h:25-69, in the reference's mixed binary32/binary64 promotions.  pow(x,2) and pow(x,5)
// are formed by multiplication in binary64 (<= 2 ulp of binary64 from glibc's pow, invisible after
// the rounding to binary32 that follows except with probability ~1e-8).
// ------------------------------------------------------------------------------------------------
// Material.h:66-69 with k = float(alpha*sqrt(2/pi)) precomputed
RT_DI float g_schlick(const DMaterial& m, float nw) {
  return __fdiv_rn(nw, __fadd_rn(__fmul_rn(nw, m.one_minus_k), m.k));
}
// The per-hit part of evaluateColorResponse: normal and outgoing direction are re-normalised by the
// reference on every call (Material.h:29-30); the result only depends on the hit, so we do it once.
struct BsdfFrame {
  float3 n, wo;   // normalize(normal), normalize(wo)
  float n_wo;     // dot(n, wo)
  float g_wo;     // gSchlick(wo, n)
};
RT_DI BsdfFrame bsdf_frame(const DMaterial& m, float3 normal, float3 wo) {
  BsdfFrame f;
  f.n = v_norm(normal);
  f.wo = v_norm(wo);
  f.n_wo = v_dot(f.n, f.wo);
  f.g_wo = g_schlick(m, f.n_wo);
  return f;
}
// Material.h:25-60 for one incoming direction wi (not yet normalised)
RT_DI float3 evaluate_color_response(const DMaterial& m, const BsdfFrame& f, float3 wi_raw) {
  const double kPi = 3.141592653589793;
  const float3 wi = v_norm(wi_raw);
  const float3 wh = v_norm(v_add(wi, f.wo));
  const double nh = (double)v_dot(f.n, wh);
  const double inner = __dadd_rn(1.0, __dmul_rn((double)m.a2m1, __dmul_rn(nh, nh)));
  const float D = (float)__ddiv_rn((double)m.a2, __dmul_rn(kPi, __dmul_rn(inner, inner)));
  const double x = __dsub_rn(1.0, fmax(0.0, (double)v_dot(wi, wh)));
  const double x2 = __dmul_rn(x, x);
  const float fr = (float)__dmul_rn(__dmul_rn(x2, x2), x);
  const float3 F = v_add(m.f0, v_scl(m.one_minus_f0, fr));
  const float n_wi = v_dot(f.n, wi);
  const float G = __fmul_rn(g_schlick(m, n_wi), f.g_wo);
  const float denom = (float)__dmul_rn(__dmul_rn(4.0, (double)n_wi), (double)f.n_wo);
  const float3 spec = v_dvs(v_scl(v_scl(F, D), G), denom);
  float3 r = v_add(m.diffuse_kd, v_scl(spec, m.one_minus_kd));
  if (r.x < 0.f) r.x = 0.f;
  if (r.y < 0.f) r.y = 0.f;
  if (r.z < 0.f) r.z = 0.f;
  return r;
}
RT_DI float3 evaluate_color_response(const DMaterial& m, float3 n, float3 wi, float3 wo) {
  return evaluate_color_response(m, bsdf_frame(m, n, wo), wi);
}
// Renderer.cpp:279-283: fmax(fmin(c,1),0) -- a NaN channel becomes 1
RT_DI float3 normalize_color(float3 c) {
  return f3(fmaxf(fminf(c.x, 1.f), 0.f), fmaxf(fminf(c.y, 1.f), 0.f), fmaxf(fminf(c.z, 1.f), 0.f));
}

// Renderer.cpp:36,42-43,274-277: hit normal and hit point from barycentrics (w*A + u*B) + v*C
// The variant with p0/e1/e2 also returns the hit triangle in the form mt_intersect takes (Ray.cpp:11: e1 = p1 - p0,
// e2 = p2 - p0, the same binary32 subtractions the traversal's triangle array holds).
RT_DI void hit_geometry(const DScene& S, const HitRec& h, float3& n, float3& P, int& mesh, float3& p0, float3& e1,
                        float3& e2) {
  int4 vi = __ldg(S.tri_vidx + h.gid);
  mesh = vi.w;
  float w = __fsub_rn(__fsub_rn(1.f, h.u), h.v);
  float3 n0 = f3(__ldg(S.nrm + vi.x)), n1 = f3(__ldg(S.nrm + vi.y)), n2 = f3(__ldg(S.nrm + vi.z));
  p0 = f3(__ldg(S.pos + vi.x));
  float3 p1 = f3(__ldg(S.pos + vi.y)), p2 = f3(__ldg(S.pos + vi.z));
  n = v_norm(v_add(v_add(v_scl(n0, w), v_scl(n1, h.u)), v_scl(n2, h.v)));
  P = v_add(v_add(v_scl(p0, w), v_scl(p1, h.u)), v_scl(p2, h.v));
  e1 = v_sub(p1, p0);
  e2 = v_sub(p2, p0);
}
RT_DI void hit_geometry(const DScene& S, const HitRec& h, float3& n, float3& P, int& mesh) {
  float3 p0, e1, e2;
  hit_geometry(S, h, n, P, mesh, p0, e1, e2);
}
// the hit point alone (same arithmetic): what the photon gather queries
RT_DI float3 hit_point(const DScene& S, const HitRec& h) {
  const int4 vi = __ldg(S.tri_vidx + h.gid);
  const float w = __fsub_rn(__fsub_rn(1.f, h.u), h.v);
  const float3 p0 = f3(__ldg(S.pos + vi.x)), p1 = f3(__ldg(S.pos + vi.y)), p2 = f3(__ldg(S.pos + vi.z));
  return v_add(v_add(v_scl(p0, w), v_scl(p1, h.u)), v_scl(p2, h.v));
}

// ------------------------------------------------------------------------------------------------
// kdtree::knearest (kdtree.h:87-107,180-195), quirks included.  The max-heap reproduces libstdc++'s
// make_heap / pop_heap / push_heap / sort_heap element moves so that ties end in the same order.
// hd/hi: this thread's heap arrays (distance, node index), element j at hd[j*hs].
// ------------------------------------------------------------------------------------------------
struct KdHeap {
  float* hd;
  int* hi;
  int hs;  // stride between consecutive heap elements
  RT_DI float d(int j) const { return hd[j * hs]; }
  RT_DI int i(int j) const { return hi[j * hs]; }
  RT_DI void set(int j, float dd, int ii) {
    hd[j * hs] = dd;
    hi[j * hs] = ii;
  }
  RT_DI void move(int dst, int src) { set(dst, d(src), i(src)); }
  // std::__push_heap
  RT_DI void push_up(int hole, int top, float vd, int vi) {
    int parent = (hole - 1) / 2;
    while (hole > top && d(parent) < vd) {
      move(hole, parent);
      hole = parent;
      parent = (hole - 1) / 2;
    }
    set(hole, vd, vi);
  }
  // std::__adjust_heap
  RT_DI void adjust(int hole, int len, float vd, int vi) {
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
      second = 2 * (second + 1);
      if (d(second) < d(second - 1)) second--;
      move(hole, second);
      hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
      second = 2 * (second + 1);
      move(hole, second - 1);
      hole = second - 1;
    }
    push_up(hole, top, vd, vi);
  }
  // std::make_heap over [0,len)
  RT_DI void make(int len) {
    if (len < 2) return;
    for (int parent = (len - 2) / 2;; parent--) {
      float vd = d(parent);
      int vi = i(parent);
      adjust(parent, len, vd, vi);
      if (parent == 0) return;
    }
  }
  // std::pop_heap over [0,len): max goes to len-1
  RT_DI void pop(int len) {
    if (len > 1) {
      float vd = d(len - 1);
      int vi = i(len - 1);
      move(len - 1, 0);
      adjust(0, len - 1, vd, vi);
    }
  }
  // std::sort_heap over [0,len)
  RT_DI void sort(int len) {
    while (len > 1) {
      pop(len);
      len--;
    }
  }
};

// Fills heap (sorted ascending on return).  kd_stack: this thread's frames, 3 ints per frame at
// stride ks: begin, end|axis<<30 packed?  -> we keep (begin, end, dx bits) and recompute the axis from depth.
RT_DI void kd_knearest(const DScene& S, float3 q, int k, KdHeap& H, int* kst, int ks, unsigned long long& visits) {
  for (int j = 0; j < k; j++) H.set(j, v_dist(f3(__ldg(S.kd_pos + j)), q), j);  // kdtree.h:186
  H.make(k);
  float best = H.d(0);  // m_bestdist (a distance, not squared)
  // explicit recursion: frame = (far_begin, far_end, dx, axis_of_children)
  int sp = 0;
  int b = 0, e = S.kd_count, axis = 0;
  for (;;) {
    // descend along the near side
    while (e > b) {
      int n = b + (e - b) / 2;
      visits++;
      float4 p = __ldg(S.kd_pos + n);
      float dnode = v_dist(f3(p), q);
      if (dnode < best) {
        H.pop(k);
        best = H.d(0);  // distance of the heap's new top BEFORE the insertion (kdtree.h:93-96)
        H.set(k - 1, dnode, n);
        H.push_up(k - 1, 0, dnode, n);
      }
      if (best == 0.f) {
        b = e;  // kdtree.h:101: return without visiting children
        break;
      }
      float pa = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
      float qa = axis == 0 ? q.x : (axis == 1 ? q.y : q.z);
      float dx = __fsub_rn(pa, qa);
      int nb, ne, fb, fe;
      if (dx > 0.f) {
        nb = b;
        ne = n;
        fb = n + 1;
        fe = e;
      } else {
        nb = n + 1;
        ne = e;
        fb = b;
        fe = n;
      }
      axis = axis == 2 ? 0 : axis + 1;
      kst[(3 * sp) * ks] = fb;
      kst[(3 * sp + 1) * ks] = fe | (axis << 28);
      kst[(3 * sp + 2) * ks] = __float_as_int(dx);
      sp++;
      b = nb;
      e = ne;
    }
    // unwind: first frame whose far side survives `dx*dx >= m_bestdist` (kdtree.h:105, squared vs not).
    // Second condition: a far subtree lies entirely beyond the splitting plane, so every photon in it has
    // d >= |dx|; when |dx| >= m_bestdist none of them can pass `d < m_bestdist` (kdtree.h:92), the visit
    // would change neither the heap nor m_bestdist, and skipping it returns exactly the reference's result
    // (only kdtree::visited() differs).  For m_bestdist < 1 -- the usual case in this scene -- this is the
    // pruning an exact k-NN would do and removes ~85 % of the node visits.  The factor keeps a 2-ulp
    // margin for sqrt(fl(a*a)) < |a|; in doubt we visit, which is what the reference does.
    bool resumed = false;
    while (sp > 0) {
      sp--;
      float dx = __int_as_float(kst[(3 * sp + 2) * ks]);
      double dx2 = __dmul_rn((double)dx, (double)dx);
      if (dx2 >= (double)best) continue;
      if (fabsf(dx) * 0.9999995f >= best) continue;
      int fe = kst[(3 * sp + 1) * ks];
      b = kst[(3 * sp) * ks];
      axis = (fe >> 28) & 3;
      e = fe & 0x0fffffff;
      resumed = true;
      break;
    }
    if (!resumed) break;
  }
  H.sort(k);
}

// ------------------------------------------------------------------------------------------------
// kd_knearest_sorted: the same query (kdtree.h:87-107,180-195, quirks included) restated for SIMT.
//
// ncu on the heap version inside k_shade (profiles/r1_photon_k_shade_full.csv): 4.9-8.8 of 32 lanes active and
// 26-32 % of the issue slots spent in the LSU on the local-memory heap and recursion stack.  What the reference
// computes only depends on the candidate SET and its two largest distances:
//     if (d < best) { evict the largest; best = largest of what is left (BEFORE the insertion); insert }
// (for k = 1 "what is left" is empty and libstdc++'s front() still returns the evicted element, kdtree.h:93-96),
// and sort_heap returns the set in ascending distance.  So the candidates live in an ascending array (an
// insertion is one predictable shift loop, the result needs no final sort), array and stack sit in shared
// memory ([slot][thread], conflict-free for any per-thread index), and the recursion is one loop in which every
// iteration visits exactly one node.
// Ties: two candidates at exactly the same binary32 distance (0.03 % of the queries at k = 50 / 357 k photons;
// also a seed node that is visited again while it is still a candidate) are ordered -- and, when they tie for
// the largest, evicted -- by libstdc++'s heap mechanics.  The function returns true when it has seen such a tie
// and the caller repeats that query with kd_knearest, the literal heap restatement; otherwise the result is
// exactly the reference's (tests: test_kdtree_and_knn_match_reference, test_knn_parity_at_scale).
// sd/si: k slots, kst: 3 ints per frame, all with stride ks between a thread's consecutive slots.
// ------------------------------------------------------------------------------------------------
// Vec3::dist (Vec3.h:216-218) is sqrt(dot(a - b, a - b)); the k-NN loops compare distances, and a correctly rounded
// square root costs ~10 instructions per visited node.  kd_dist2 is the argument of that square root (same operation
// order), and kd_reject_from(best) a bound B with:  d2 >= B  =>  sqrt_rn(d2) >= best  (B = best*best rounded UP is
// >= best^2, sqrt is monotone and sqrt_rn(best^2) = best).  So `dist < best` is false without taking the root
// whenever d2 >= B; only the candidates that can actually enter pay for the root and the exact comparison.
RT_DI float kd_dist2(float3 p, float3 q) {
  const float3 a = v_sub(p, q);
  return v_dot(a, a);
}
RT_DI float kd_reject_from(float best) { return __fmul_ru(best, best); }

RT_DI unsigned long long kd_pack(float d, int idx) {  // distances are >= 0: their bit patterns order like the values
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
}
RT_DI float kd_dist_of(unsigned long long c) { return __uint_as_float((unsigned)(c >> 32)); }
RT_DI int kd_index_of(unsigned long long c) { return (int)(unsigned)c; }

// ------------------------------------------------------------------------------------------------
// kd_knearest_heap: the same single-loop traversal with the candidates kept as libstdc++'s own max-heap (the element
// moves of make_heap / pop_heap / push_heap / sort_heap restated literally, std::__adjust_heap and std::__push_heap)
// over packed words in shared memory.  A new candidate is usually small, so the eviction costs one sift-down of
// ~log2(k) levels and the insertion stops at the leaf, against ~0.6 k shifts in the ascending array; ties need no
// special case because the moves ARE the reference's.  Selected with -DRT_KNN_HEAP=1 (see profiles/r1_tuning.md).
// ------------------------------------------------------------------------------------------------
struct KdHeapP {
  unsigned long long* a;
  int hs;
  RT_DI unsigned long long get(int j) const { return a[(unsigned)j * (unsigned)hs]; }
  RT_DI void set(int j, unsigned long long w) { a[(unsigned)j * (unsigned)hs] = w; }
  RT_DI static bool less(unsigned long long x, unsigned long long y) { return (unsigned)(x >> 32) < (unsigned)(y >> 32); }
  RT_DI void push_up(int hole, int top, unsigned long long v) {  // std::__push_heap
    int parent = (hole - 1) / 2;
    while (hole > top) {
      const unsigned long long pw = get(parent);
      if (!less(pw, v)) break;
      set(hole, pw);
      hole = parent;
      parent = (hole - 1) / 2;
    }
    set(hole, v);
  }
  RT_DI void adjust(int hole, int len, unsigned long long v) {  // std::__adjust_heap
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
      second = 2 * (second + 1);
      unsigned long long ws = get(second);
      const unsigned long long wl = get(second - 1);
      if (less(ws, wl)) {
        second--;
        ws = wl;
      }
      set(hole, ws);
      hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
      second = 2 * (second + 1);
      set(hole, get(second - 1));
      hole = second - 1;
    }
    push_up(hole, top, v);
  }
  RT_DI void make(int len) {  // std::make_heap
    if (len < 2) return;
    for (int parent = (len - 2) / 2;; parent--) {
      adjust(parent, len, get(parent));
      if (parent == 0) return;
    }
  }
  RT_DI void pop(int len) {  // std::pop_heap: the largest goes to slot len-1
    if (len > 1) {
      const unsigned long long v = get(len - 1);
      set(len - 1, get(0));
      adjust(0, len - 1, v);
    }
  }
  RT_DI void sort(int len) {  // std::sort_heap
    while (len > 1) {
      pop(len);
      len--;
    }
  }
};
// ------------------------------------------------------------------------------------------------
// KdQuery: ONE kdtree::knearest query as a resumable state machine -- init() seeds the candidates, every step()
// visits exactly one node and unwinds the stack to the next one, finish() leaves the k candidates in ascending
// distance.  The persistent gather kernel (k_knn_gather) interleaves the steps of 32 queries per warp and refills a
// lane as soon as its query completes; kd_knearest_sorted / kd_knearest_heap below run one query to completion.
//   POLICY 0: ascending candidate array (small k): an insertion is one backward shift loop; ties are reported (see
//             above) and the query is repeated with the literal heap.
//   POLICY 1: libstdc++'s own max-heap moves restated (KdHeapP): log2(k) moves per eviction, ties for free.
// sc: k candidate slots (distance bits << 32 | node index, one 64-bit access per move), stride cs between a thread's
// slots; kst: 3 ints per frame (far begin, far end | next axis << 28, threshold bits), stride ks.
// ------------------------------------------------------------------------------------------------
#ifndef RT_KNN_FRONT_SCAN
#define RT_KNN_FRONT_SCAN 0
#endif
template <int POLICY>
struct KdQuery {
  float3 q;
  float best;     // m_bestdist (a distance, not squared)
  float reject2;  // d2 >= reject2  =>  sqrt_rn(d2) >= best
  int sp, b, e, axis;
  unsigned nv;
  bool tie;

  RT_DI void init(const DScene& S, float3 query, int k, unsigned long long* sc, int cs) {
    q = query;
    tie = false;
    if (POLICY == 0) {
      for (int j = 0; j < k; j++) {  // kdtree.h:186: the first k nodes of the array seed the candidates
        const unsigned dj = __float_as_uint(v_dist(f3(__ldg(S.kd_pos + j)), q));
        int m = j - 1;
        unsigned long long w = 0;
        while (m >= 0 && (unsigned)((w = sc[(unsigned)(m) * (unsigned)cs]) >> 32) > dj) {
          sc[(unsigned)(m + 1) * (unsigned)cs] = w;
          m--;
        }
        if (m >= 0 && (unsigned)(w >> 32) == dj) tie = true;
        sc[(unsigned)(m + 1) * (unsigned)cs] = ((unsigned long long)dj << 32) | (unsigned)j;
      }
      best = kd_dist_of(sc[(unsigned)(k - 1) * (unsigned)cs]);
    } else {
      KdHeapP H{sc, cs};
      for (int j = 0; j < k; j++) H.set(j, kd_pack(v_dist(f3(__ldg(S.kd_pos + j)), q), j));
      H.make(k);
      best = kd_dist_of(H.get(0));
    }
    reject2 = kd_reject_from(best);
    sp = 0;
    b = 0;
    e = S.kd_count;
    axis = 0;
    nv = 0;
  }

  // one node visit (precondition: e > b); returns false when the traversal is complete
  RT_DI bool step(const DScene& S, int k, unsigned long long* sc, int cs, int* kst, int ks) {
    const int n = b + (e - b) / 2;
    nv++;
    const float4 p = __ldg(S.kd_pos + n);
    const float d2 = kd_dist2(f3(p), q);
    float dnode;
    if (d2 < reject2 && (dnode = __fsqrt_rn(d2)) < best) {  // kdtree.h:92-99
      if (POLICY == 0) {
        // evict the largest; m_bestdist = the largest of the rest BEFORE the insertion (k == 1: libstdc++'s
        // front() after pop_heap is the evicted candidate itself), kdtree.h:93-96
        best = kd_dist_of(sc[(unsigned)(k > 1 ? k - 2 : 0) * (unsigned)cs]);
        const unsigned dn = __float_as_uint(dnode);
#if RT_KNN_FRONT_SCAN
        // (measured on B200, cfg3 frame: 63.8 ms against 57.9 for the single backward loop below: the new candidate
        // does land ~2 slots from the front, but two loops diverge twice)
        int pos = 0;  // first slot of [0, k-1) whose distance is above dn: the new candidate goes there
        while (pos < k - 1) {
          const unsigned dw = (unsigned)(sc[(unsigned)pos * (unsigned)cs] >> 32);
          if (dw > dn) break;
          if (dw == dn) tie = true;
          pos++;
        }
        for (int j = k - 2; j >= pos; j--) sc[(unsigned)(j + 1) * (unsigned)cs] = sc[(unsigned)j * (unsigned)cs];
        sc[(unsigned)pos * (unsigned)cs] = ((unsigned long long)dn << 32) | (unsigned)n;
#else
        int m = k - 2;
        unsigned long long w = 0;
        while (m >= 0 && (unsigned)((w = sc[(unsigned)(m) * (unsigned)cs]) >> 32) > dn) {
          sc[(unsigned)(m + 1) * (unsigned)cs] = w;
          m--;
        }
        // a tie with the evicted largest (old slot k-1) cannot matter: dnode < best <= it
        if (m >= 0 && (unsigned)(w >> 32) == dn) tie = true;
        sc[(unsigned)(m + 1) * (unsigned)cs] = ((unsigned long long)dn << 32) | (unsigned)n;
#endif
      } else {
        KdHeapP H{sc, cs};
        H.pop(k);
        best = kd_dist_of(H.get(0));  // the new top BEFORE the insertion (for k == 1: the evicted candidate itself)
        H.push_up(k - 1, 0, kd_pack(dnode, n));
      }
      reject2 = kd_reject_from(best);
    }
    int nb = b, ne = b;  // empty
    if (best != 0.f) {   // kdtree.h:101: best == 0 returns without visiting the children
      float pa = p.x, qa = q.x;
      if (axis == 1) pa = p.y, qa = q.y;
      if (axis == 2) pa = p.z, qa = q.z;
      const float dx = __fsub_rn(pa, qa);
      const bool left_near = dx > 0.f;
      nb = left_near ? b : n + 1;
      ne = left_near ? n : e;
      const int fb = left_near ? n + 1 : b, fe = left_near ? e : n;
      axis = axis == 2 ? 0 : axis + 1;
      if (fe > fb) {  // an empty far side has nothing to visit
        // The far side is skipped when `dx*dx >= m_bestdist` (kdtree.h:105: a square against a distance, in
        // binary64, where the square of a binary32 is exact) or when |dx| >= m_bestdist (every photon behind the
        // plane is at least |dx| away, none can pass `d < m_bestdist`, so the visit would change nothing; the
        // factor keeps a 2-ulp margin for sqrt(fl(a*a)) < |a|).  Both are "m_bestdist <= a number known now":
        // the largest binary32 <= dx*dx, and fl(|dx| * 0.9999995) -- so one threshold is stored and the pop
        // costs one comparison.
        // For k >= 2 m_bestdist never grows (it is the second largest candidate and candidates only get
        // closer), so a far side that is already skippable now is not pushed at all.  (k == 1: m_bestdist is
        // the distance of the candidate evicted LAST, which can go up again.)
        const float t_sq = __fmul_rd(dx, dx);  // the largest binary32 <= dx*dx (the binary64 square is exact)
        const float t_pl = __fmul_rn(fabsf(dx), 0.9999995f);
        const float thr = fmaxf(t_sq, t_pl);
        if (best > thr || k == 1) {
          kst[(3 * sp) * ks] = fb;
          kst[(3 * sp + 1) * ks] = fe | (axis << 28);
          kst[(3 * sp + 2) * ks] = __float_as_int(thr);
          sp++;
        }
      }
    }
    b = nb;
    e = ne;
    while (e <= b && sp > 0) {  // near side empty: resume at the newest frame whose far side survives
      sp--;
      if (best <= __int_as_float(kst[(3 * sp + 2) * ks])) continue;
      const int fe = kst[(3 * sp + 1) * ks];
      b = kst[(3 * sp) * ks];
      axis = (fe >> 28) & 3;
      e = fe & 0x0fffffff;
    }
    return e > b;
  }

  RT_DI void finish(int k, unsigned long long* sc, int cs) {
    if (POLICY == 1) {
      KdHeapP H{sc, cs};
      H.sort(k);
    }
  }
};

// one query run to completion; the sorted flavour returns true when it saw a tie (the caller repeats the query with
// kd_knearest, the literal heap restatement)
RT_DI bool kd_knearest_sorted(const DScene& S, float3 q, int k, unsigned long long* sc, int cs, int* kst, int ks,
                              unsigned long long& visits) {
  KdQuery<0> Q;
  Q.init(S, q, k, sc, cs);
  while (Q.step(S, k, sc, cs, kst, ks)) {
  }
  visits += Q.nv;
  return Q.tie;
}
RT_DI void kd_knearest_heap(const DScene& S, float3 q, int k, unsigned long long* sc, int cs, int* kst, int ks,
                            unsigned long long& visits) {
  KdQuery<1> Q;
  Q.init(S, q, k, sc, cs);
  while (Q.step(S, k, sc, cs, kst, ks)) {
  }
  Q.finish(k, sc, cs);
  visits += Q.nv;
}

// ------------------------------------------------------------------------------------------------
// kd_knearest_exact (RT_FLAG_KNN_EXACT; off by default because it changes 0.5-2.8 % of the
// queries against the reference): the k photons with the smallest (distance, array index), in that order -- a
// canonical exact k-NN on the same tree.  No seeds, no `best == 0` shortcut, and the bound is on the current k-th
// distance itself: a far side is skipped when every photon behind the plane is strictly farther than it.
// Candidates are the same packed (distance bits << 32 | index) words, whose integer order IS (distance, index).
// ------------------------------------------------------------------------------------------------
RT_DI void kd_knearest_exact(const DScene& S, float3 q, int k, unsigned long long* sc, int cs, int* kst, int ks,
                             unsigned long long& visits) {
  int cnt = 0;
  unsigned long long worst = ~0ull;  // packed k-th candidate; all-ones while fewer than k are held
  float reject2 = INFINITY;          // d2 >= reject2  =>  the distance is above the k-th one (see kd_reject_from)
  int sp = 0, b = 0, e = S.kd_count, axis = 0;
  unsigned nv = 0;
  while (e > b) {
    const int n = b + (e - b) / 2;
    nv++;
    const float4 p = __ldg(S.kd_pos + n);
    const float d2 = kd_dist2(f3(p), q);
    unsigned long long cand = ~0ull;
    if (!(d2 >= reject2)) cand = kd_pack(__fsqrt_rn(d2), n);  // (a NaN distance goes the slow way, as before)
    if (cand < worst) {
      int m = (cnt < k ? cnt : k - 1) - 1;  // last slot that stays
      unsigned long long w = 0;
      while (m >= 0 && (w = sc[(unsigned)(m) * (unsigned)cs]) > cand) {
        sc[(unsigned)(m + 1) * (unsigned)cs] = w;
        m--;
      }
      sc[(unsigned)(m + 1) * (unsigned)cs] = cand;
      if (cnt < k) cnt++;
      if (cnt == k) {
        worst = sc[(unsigned)(k - 1) * (unsigned)cs];
        // a tie with the k-th distance can still enter on its index: reject only from the next binary32 up
        reject2 = kd_reject_from(__uint_as_float(__float_as_uint(kd_dist_of(worst)) + 1u));
      }
    }
    float pa = p.x, qa = q.x;
    if (axis == 1) pa = p.y, qa = q.y;
    if (axis == 2) pa = p.z, qa = q.z;
    const float dx = __fsub_rn(pa, qa);
    const bool left_near = dx > 0.f;
    const int nb = left_near ? b : n + 1, ne = left_near ? n : e;
    const int fb = left_near ? n + 1 : b, fe = left_near ? e : n;
    axis = axis == 2 ? 0 : axis + 1;
    if (fe > fb) {
      // every photon in the far side is at least |dx| away (2-ulp margin for sqrt(fl(a*a)) < |a|); the k-th
      // distance only shrinks, so a side that cannot matter now is not pushed at all
      const float thr = __fmul_rn(fabsf(dx), 0.9999995f);
      if (cnt < k || !(thr > kd_dist_of(worst))) {
        kst[(3 * sp) * ks] = fb;
        kst[(3 * sp + 1) * ks] = fe | (axis << 28);
        kst[(3 * sp + 2) * ks] = __float_as_int(thr);
        sp++;
      }
    }
    b = nb;
    e = ne;
    while (e <= b && sp > 0) {
      sp--;
      if (cnt == k && __int_as_float(kst[(3 * sp + 2) * ks]) > kd_dist_of(worst)) continue;
      const int fe2 = kst[(3 * sp + 1) * ks];
      b = kst[(3 * sp) * ks];
      axis = (fe2 >> 28) & 3;
      e = fe2 & 0x0fffffff;
    }
  }
  visits += nv;
}

}  // namespace rtb
