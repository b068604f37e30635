// port.h -- CPU restatement of the render hot path of nikitakaraevv/ray-tracing-engine.
//
// TEST INFRASTRUCTURE (oracle/): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use this code; the product (ray-tracing-engine_b200/) never links it.
//
// PARITY PINNING: the reference ships no tests and no golden vectors, so this
// restatement is pinned against the reference ITSELF: tests/test_oracle.py runs every
// function below against the unmodified reference compiled into oracle/_ref/libref_cb.so (bit-exact
// comparisons), and tests/golden/ holds vectors generated from that library by oracle/gen_golden.py
// for the machines where /root/reference is absent.
//
// Everything is plain scalar C++ over flat arrays, one function per reference routine, each citing
// the reference file:line it follows.  Floating point is strict IEEE binary32/binary64 in the
// reference's operation order and promotions (build: -ffp-contract=off, no -ffast-math).
#ifndef RT_ORACLE_PORT_H
#define RT_ORACLE_PORT_H
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cfloat>
#include <vector>

#include "../rng_contract.h"

namespace orc {

struct V3 {
  float x, y, z;
};
static inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }  // Vec3.h:86-92
static inline V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }  // Vec3.h:94-100
static inline V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }                       // Vec3.h:102-108
static inline V3 mul(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }  // Vec3.h:110-116
static inline V3 scl(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }     // Vec3.h:118-124, 300-302
static inline V3 dvs(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }     // Vec3.h:134-140
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // Vec3.h:220-223
static inline V3 cross(V3 a, V3 b) {                                               // Vec3.h:225-232
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Vec3.h:165-167: length() = (T)sqrt(squaredLength()) -- the double sqrt of a float, rounded back.
static inline float length(V3 a) { return (float)sqrt((double)dot(a, a)); }
static inline float dist(V3 a, V3 b) { return length(sub(a, b)); }  // Vec3.h:215-218
// Vec3.h:170-178, 234-239: zero vectors are returned unchanged; otherwise multiply by 1/len.
static inline V3 normalize(V3 a) {
  float l = length(a);
  if (l == 0.0f) return a;
  float invL = 1.0f / l;
  return mk(a.x * invL, a.y * invL, a.z * invL);
}
// Vec3.h:180-199
static inline void two_orthogonals(V3 n, V3& u, V3& v) {
  if (fabs(n.x) < fabs(n.y)) {
    if (fabs(n.x) < fabs(n.z))
      u = mk(0, -n.z, n.y);
    else
      u = mk(-n.y, n.x, 0);
  } else {
    if (fabs(n.y) < fabs(n.z))
      u = mk(n.z, 0, -n.x);
    else
      u = mk(-n.y, n.x, 0);
  }
  v = cross(n, u);
}

// ------------------------------------------------------------------ random stream (rng_contract.h)
struct Rng {
  uint64_t key;
  uint32_t ctr;
  uint64_t drawn;
  Rng(uint64_t seed, uint64_t domain, uint64_t index) : key(rto_stream_key(seed, domain, index)), ctr(0), drawn(0) {}
  uint32_t word() {
    drawn++;
    return rto_word(key, ctr++);
  }
  // libstdc++ generate_canonical<float,24> with a 32-bit engine: one word.
  float canonical_f() {
    float f = (float)word() / 4294967296.0f;
    if (f >= 1.0f) f = nextafterf(1.0f, 0.0f);
    return f;
  }
  // libstdc++ generate_canonical<double,53> with a 32-bit engine: two words, low word first.
  double canonical_d() {
    double lo = (double)word();
    double hi = (double)word();
    double g = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    if (g >= 1.0) g = nextafter(1.0, 0.0);
    return g;
  }
  float uniform_f(float a, float b) { return canonical_f() * (b - a) + a; }
  double uniform_d(double a, double b) { return canonical_d() * (b - a) + a; }
};

// ------------------------------------------------------------------ flat scene
struct Material {  // Material.h:62-64
  float kd, alpha;
  V3 albedo, F0;
};
struct Light {  // LightSource.h:61-65 (bases are host-computed, LightSource.h:29-32)
  V3 position, color, normal, vertical, horizontal;
  float intensity, side, ac, al, aq, factor;
};
struct Camera {  // Camera.h:34-40
  V3 position, lowerLeft, horizontal, vertical;
};
struct Scene {
  int V = 0, T = 0, M = 0, L = 0, w = 0, h = 0;
  std::vector<V3> P, N;                  // per-vertex, all meshes concatenated in scene order
  std::vector<int32_t> tri;              // 3*T GLOBAL vertex indices, scene order
  std::vector<int32_t> mesh_tri_off;     // M+1
  std::vector<int32_t> mesh_vtx_off;     // M+1
  std::vector<int32_t> tri_mesh;         // T
  std::vector<Material> mats;
  std::vector<Light> lights;
  Camera cam;
};

struct Hit {
  bool found;
  int32_t mesh, tri;  // tri = global triangle index in scene order
  float u, v, d;
};

// Ray.cpp:9-24 (Moller-Trumbore, absolute epsilon on det).  u, v, t are written before the range
// tests, exactly like the reference, so callers that want them for misses get them.
static inline bool triangle_intersect(V3 o, V3 dir, V3 p0, V3 p1, V3 p2, float& u, float& v, float& t) {
  V3 edge1 = sub(p1, p0), edge2 = sub(p2, p0);
  V3 pvec = cross(dir, edge2);
  float det = dot(edge1, pvec);
  if (fabs(det) < 0.000001f) return false;
  float inv_det = 1.0f / det;
  V3 tvec = sub(o, p0);
  u = dot(tvec, pvec) * inv_det;
  V3 qvec = cross(tvec, edge1);
  v = dot(dir, qvec) * inv_det;
  t = dot(edge2, qvec) * inv_det;
  if (u < 0.f || u > 1.f) return false;
  if (v >= 0.f && u + v <= 1.f) return true;
  return false;
}

// RayTracer.h:27-53: brute force over meshes (scene order) x triangles (file order); strict
// `dt < closest` so the first index wins exact ties.
static inline Hit ray_trace(const Scene& s, V3 o, V3 dir, uint64_t* ray_counter = nullptr) {
  if (ray_counter) ++*ray_counter;
  Hit h{false, 0, 0, 0.f, 0.f, 0.f};
  float closest = FLT_MAX;
  for (int t = 0; t < s.T; t++) {
    float ut, vt, dt;
    if (triangle_intersect(o, dir, s.P[s.tri[3 * t]], s.P[s.tri[3 * t + 1]], s.P[s.tri[3 * t + 2]], ut, vt, dt)) {
      if (dt > 0.f && dt < closest) {
        h.found = true;
        closest = dt;
        h.tri = t;
        h.mesh = s.tri_mesh[t];
        h.u = ut;
        h.v = vt;
        h.d = dt;
      }
    }
  }
  return h;
}

// RayTracer.h:109-117.  The sums are formed in double (float + double) and rounded once.
static inline void jitter_sample(Rng& g, int sampleIdx, int nSamples, float& x, float& y) {
  int d = (int)sqrtf((float)nSamples);
  int j2 = sampleIdx / d;
  int i2 = sampleIdx % d;
  x = (float)(((double)(float)i2 + g.uniform_d(0.0, 1.0)) / (double)(float)d);
  y = (float)(((double)(float)j2 + g.uniform_d(0.0, 1.0)) / (double)(float)d);
}

// RayTracer.h:95-107 with maxRayAngle = float(M_PI/2): the distribution's upper bound is
// 2*1.57079637f/pi = 1.0000000278 (so asin can return NaN with p ~ 2.8e-8, kept).
static inline V3 hsphere_uniform_sample(Rng& g, V3 normal) {
  const float maxRayAngle = (float)(M_PI / 2.f);
  const double hi = (double)(2 * maxRayAngle) / M_PI;
  normal = normalize(normal);
  V3 v1, v2;
  two_orthogonals(normal, v1, v2);
  v1 = normalize(v1);
  v2 = normalize(v2);
  float theta = (float)asin(g.uniform_d(0.0, hi));
  float phi = (float)(2 * M_PI * g.uniform_d(0.0, hi));
  V3 direction = add(scl(v1, cosf(phi)), scl(v2, sinf(phi)));
  direction = normalize(direction);
  return normalize(add(scl(normal, cosf(theta)), scl(direction, sinf(theta))));
}

// Camera.h:27-30
static inline void camera_ray(const Camera& c, float u, float v, V3& o, V3& d) {
  o = c.position;
  d = normalize(sub(add(add(c.lowerLeft, scl(c.horizontal, u)), scl(c.vertical, v)), c.position));
}

// LightSource.h:46-49.  g++ 13 evaluates the right-hand operand's draw first, so the FIRST draw
// scales m_horizontal and the SECOND m_vertical (verified against the reference build).
static inline V3 light_rand_area_position(const Light& l, Rng& g) {
  float a = g.uniform_f(-l.side, l.side);  // -> horizontal
  float b = g.uniform_f(-l.side, l.side);  // -> vertical
  return add(add(l.position, scl(l.vertical, b)), scl(l.horizontal, a));
}
// LightSource.h:51-54
static inline float light_radiance(const Light& l, V3 p) {
  float d = dist(p, l.position);
  return l.intensity / (l.ac + l.al * d + l.aq * d * d);
}
// LightSource.h:56-59
static inline V3 light_evaluate(const Light& l, V3 p) { return scl(scl(l.color, l.factor), light_radiance(l, p)); }

// Material.h:66-69
static inline float g_schlick(const Material& m, V3 w, V3 n) {
  float k = (float)(m.alpha * sqrt(2. / M_PI));
  return dot(n, w) / (dot(n, w) * (1 - k) + k);
}
// Material.h:41-60 (mixed float/double exactly as the reference's promotions resolve)
static inline V3 specular_response(const Material& m, V3 n, V3 wi, V3 wo) {
  V3 wh = normalize(add(wi, wo));
  float a2 = m.alpha * m.alpha;
  float D = (float)(a2 / (M_PI * pow(1 + (a2 - 1) * pow((double)dot(n, wh), 2), 2)));
  float fr = (float)pow(1 - fmax((double)0, (double)dot(wi, wh)), 5);
  V3 F = add(m.F0, scl(sub(mk(1.f, 1.f, 1.f), m.F0), fr));
  float G = g_schlick(m, wi, n) * g_schlick(m, wo, n);
  float denom = (float)(4. * dot(n, wi) * dot(n, wo));
  return dvs(scl(scl(F, D), G), denom);
}
// Material.h:25-39
static inline V3 evaluate_color_response(const Material& m, V3 n, V3 wi, V3 wo) {
  V3 diffuse = dvs(m.albedo, (float)M_PI);
  V3 r = add(scl(diffuse, m.kd), scl(specular_response(m, normalize(n), normalize(wi), normalize(wo)), 1 - m.kd));
  if (r.x < 0.f) r.x = 0.f;
  if (r.y < 0.f) r.y = 0.f;
  if (r.z < 0.f) r.z = 0.f;
  return r;
}

// Renderer.cpp:274-277 with w = 1 - u - v (Renderer.cpp:36)
static inline V3 bary(const std::vector<V3>& a, const int32_t* t, float w, float u, float v) {
  return add(add(scl(a[t[0]], w), scl(a[t[1]], u)), scl(a[t[2]], v));
}

// ------------------------------------------------------------------ photon map (PhotonMap.h, kdtree.h)
struct Particle {  // Particle.h:33-35
  V3 position, direction;
  float weight;
};
struct KdNode {
  Particle p;
  int32_t left, right;
};
struct KdTree {
  std::vector<KdNode> nodes;
  int32_t root = -1;
  uint64_t visited = 0;
  // kdtree.h:60-69: in-place median tree, std::nth_element decides tie placement (same libstdc++).
  int32_t make_tree(size_t begin, size_t end, size_t index) {
    if (end <= begin) return -1;
    size_t n = begin + (end - begin) / 2;
    std::nth_element(nodes.begin() + begin, nodes.begin() + n, nodes.begin() + end,
                     [index](const KdNode& a, const KdNode& b) {
                       const float* pa = &a.p.position.x;
                       const float* pb = &b.p.position.x;
                       return pa[index] < pb[index];
                     });
    index = (index + 1) % 3;
    nodes[n].left = make_tree(begin, n, index);
    nodes[n].right = make_tree(n + 1, end, index);
    return (int32_t)n;
  }
  void build(const std::vector<Particle>& list) {  // kdtree.h:119-126
    nodes.clear();
    nodes.reserve(list.size());
    for (const Particle& p : list) nodes.push_back(KdNode{p, -1, -1});
    root = make_tree(0, nodes.size(), 0);
  }
  bool empty() const { return nodes.empty(); }

  struct HeapItem {
    double d;
    int32_t node;
  };
  // kdtree.h:87-107 -- quirks kept: m_bestdist becomes the distance of the heap's NEW top after the
  // pop (i.e. before the candidate is inserted), and the far-side prune compares dx*dx (squared)
  // with m_bestdist (not squared).
  void knearest_rec(int32_t node, V3 q, size_t index, std::vector<HeapItem>& heap, double& bestdist) {
    if (node < 0) return;
    ++visited;
    const KdNode& nd = nodes[node];
    double d = (double)dist(nd.p.position, q);
    auto cmp = [](const HeapItem& a, const HeapItem& b) { return a.d < b.d; };
    if (d < bestdist) {
      std::pop_heap(heap.begin(), heap.end(), cmp);
      HeapItem top = heap.front();
      heap.pop_back();
      bestdist = top.d;
      heap.push_back(HeapItem{d, node});
      std::push_heap(heap.begin(), heap.end(), cmp);
    }
    if (bestdist == 0) return;
    const float* pp = &nd.p.position.x;
    const float* qq = &q.x;
    double dx = (double)(pp[index] - qq[index]);
    index = (index + 1) % 3;
    knearest_rec(dx > 0 ? nd.left : nd.right, q, index, heap, bestdist);
    if (dx * dx >= bestdist) return;
    knearest_rec(dx > 0 ? nd.right : nd.left, q, index, heap, bestdist);
  }
  // kdtree.h:180-195.  Returns node indices in the reference's output order (ascending distance).
  // Precondition (the reference throws otherwise): !empty() and k <= nodes.size().
  void knearest(V3 q, int k, std::vector<int32_t>& out) {
    std::vector<HeapItem> heap;
    for (int i = 0; i < k; i++) heap.push_back(HeapItem{(double)dist(nodes[i].p.position, q), i});
    auto cmp = [](const HeapItem& a, const HeapItem& b) { return a.d < b.d; };
    std::make_heap(heap.begin(), heap.end(), cmp);
    visited = 0;
    double bestdist = heap[0].d;
    knearest_rec(root, q, 0, heap, bestdist);
    std::sort_heap(heap.begin(), heap.end(), cmp);
    out.clear();
    for (int i = 0; i < k; i++) out.push_back(heap[i].node);
  }
};

struct Counters {
  uint64_t rays = 0;     // RayTracer::rayTrace invocations (primary, bounce, shadow, photon segments)
  uint64_t queries = 0;  // kdtree::knearest invocations
  uint64_t visits = 0;   // kd nodes visited
};

// Renderer.cpp:33-61 (direct lighting).  Every light draws its 2 uniforms before the occlusion test.
static inline V3 shade_direct(const Scene& s, V3 rayDir, const Hit& h, Rng& g, V3& hitNormal, V3& P, Counters& c) {
  float w = 1.f - h.u - h.v;
  const int32_t* t = &s.tri[3 * h.tri];
  hitNormal = normalize(bary(s.N, t, w, h.u, h.v));
  P = bary(s.P, t, w, h.u, h.v);
  const Material& m = s.mats[h.mesh];
  V3 color = mk(0.f, 0.f, 0.f);
  for (const Light& l : s.lights) {
    V3 toLight = sub(light_rand_area_position(l, g), P);
    if (ray_trace(s, P, toLight, &c.rays).found) continue;
    V3 bsdf = evaluate_color_response(m, hitNormal, toLight, neg(rayDir));
    V3 radiance = light_evaluate(l, P);
    color = add(color, mul(radiance, bsdf));
  }
  return color;
}

// Renderer.cpp:63-104 (photon gather).  numPhotons is the REQUESTED count; factor 100 (Renderer.h:45).
static inline V3 shade_photon(const Scene& s, V3 rayDir, const Hit& h, KdTree& tree, int k, int numPhotons,
                              V3& hitNormal, V3& P, Counters& c) {
  float w = 1.f - h.u - h.v;
  const int32_t* t = &s.tri[3 * h.tri];
  hitNormal = normalize(bary(s.N, t, w, h.u, h.v));
  P = bary(s.P, t, w, h.u, h.v);
  const Material& m = s.mats[h.mesh];
  std::vector<int32_t> result;
  tree.knearest(P, k, result);
  c.queries++;
  c.visits += tree.visited;
  float r = dist(tree.nodes[result[k - 1]].p.position, P);
  float area = (float)(M_PI * r * r);
  V3 averageDirection = mk(0.f, 0.f, 0.f), radiance = mk(0.f, 0.f, 0.f);
  for (int32_t idx : result) {
    averageDirection = add(averageDirection, tree.nodes[idx].p.direction);
    radiance = add(radiance, mk(1.f, 1.f, 1.f));
  }
  radiance = dvs(radiance, area);
  radiance = dvs(radiance, (float)numPhotons);
  radiance = scl(radiance, 100.f);
  V3 bsdf = evaluate_color_response(m, hitNormal, normalize(averageDirection), neg(rayDir));
  return mul(radiance, bsdf);
}

// Renderer.cpp:106-201.  mode 0 = calculateColorRay, mode 1 = calculateColorPath (finalDepth 3);
// tree == nullptr or empty selects the direct-lighting overloads (Renderer.cpp:237-250).
static inline V3 calculate_color(const Scene& s, V3 o, V3 d, int mode, KdTree* tree, int k, int numPhotons, Rng& g,
                                 bool& posIntersectionFound, Counters& c, int depth = 0) {
  const int finalDepth = 3;
  V3 zero = mk(0.f, 0.f, 0.f);
  if (mode == 1 && depth >= finalDepth) return dvs(zero, (float)depth);
  Hit h = ray_trace(s, o, d, &c.rays);
  if (!(h.found && h.d > 0.f)) {
    if (mode == 0 || depth == 0) posIntersectionFound = false;
    return zero;
  }
  V3 n, P, color;
  if (tree && !tree->empty())
    color = shade_photon(s, d, h, *tree, k, numPhotons, n, P, c);
  else
    color = shade_direct(s, d, h, g, n, P, c);
  if (mode == 0) return color;
  V3 randomDirection = hsphere_uniform_sample(g, n);
  return add(color, calculate_color(s, P, randomDirection, mode, tree, k, numPhotons, g, posIntersectionFound, c,
                                    depth + 1));
}

// Renderer.cpp:279-283 (fmin/fmax: a NaN channel becomes 1)
static inline V3 normalize_color(V3 c) {
  return mk(fmaxf(fminf(c.x, 1.f), 0.f), fmaxf(fminf(c.y, 1.f), 0.f), fmaxf(fminf(c.z, 1.f), 0.f));
}

// One pixel sample: Renderer.cpp:228-258.
static inline V3 render_sample(const Scene& s, int N, int mode, KdTree* tree, int k, int numPhotons, uint64_t seed,
                               int x, int y, int i, bool& found, Counters& c) {
  Rng g(seed, RTO_DOMAIN_PIXEL, (uint64_t)i * ((uint64_t)s.w * s.h) + (uint64_t)y * s.w + (uint64_t)x);
  float shiftX, shiftY;
  jitter_sample(g, i, N, shiftX, shiftY);
  V3 o, d;
  camera_ray(s.cam, ((float)x + shiftX) / (float)s.w, 1.f - ((float)y + shiftY) / (float)s.h, o, d);
  found = true;
  V3 color = calculate_color(s, o, d, mode, tree, k, numPhotons, g, found, c);
  return normalize_color(color);
}

// PhotonMap.h:92-155, one path (iterative form of the tail recursion).  Returns the reference's
// return value (depth of the Russian-roulette kill, or -1) and appends at most one particle.
static inline int photon_path(const Scene& s, V3 o, V3 d, Particle photon, Rng& g, std::vector<Particle>& list,
                              Counters& c) {
  const int max_depth = 20;
  bool exit = false;
  for (int depth = 0;; depth++) {
    if (exit) {
      list.push_back(photon);
      return depth - 1;
    }
    if (depth >= max_depth) return -1;
    Hit h = ray_trace(s, o, d, &c.rays);
    if (!(h.found && h.d > 0.f)) {
      if (depth != 0) list.push_back(photon);
      return -1;
    }
    float w = 1.f - h.u - h.v;
    const int32_t* t = &s.tri[3 * h.tri];
    const Material& m = s.mats[h.mesh];
    V3 hitNormal = normalize(bary(s.N, t, w, h.u, h.v));
    V3 P = bary(s.P, t, w, h.u, h.v);
    photon.position = P;
    photon.direction = neg(d);
    V3 randomDirection = hsphere_uniform_sample(g, hitNormal);
    V3 perfectReflection = sub(d, scl(hitNormal, 2.f * dot(d, hitNormal)));
    float bsdf = length(evaluate_color_response(m, hitNormal, d, randomDirection));
    float pdf = (dot(normalize(randomDirection), normalize(perfectReflection)) + 1.f) / 2.f;
    photon.weight *= bsdf / pdf;
    float continueProb = fminf(photon.weight, 1.f);
    if (g.uniform_f(0.f, 1.f) > continueProb)
      exit = true;
    else
      photon.weight /= continueProb;
    o = P;
    d = randomDirection;
  }
}

// PhotonMap.h:14-50: paths [p0,p1) of every light; path p of light l draws from stream
// (seed, PHOTON, l*perLight + p).
static inline int photons_per_light(const Scene& s, int numPhotons) {
  float lightPdf = 1.f / (float)s.lights.size();
  return (int)((float)numPhotons * lightPdf);
}
static inline void emit_photons(const Scene& s, int numPhotons, uint64_t seed, int p0, int p1,
                                std::vector<Particle>& list, int* depth_hist, Counters& c,
                                std::vector<int64_t>* per_light = nullptr) {
  if (numPhotons <= 0) return;
  float lightPdf = 1.f / (float)s.lights.size();
  int perLight = photons_per_light(s, numPhotons);
  if (p0 < 0) p0 = 0;
  if (p1 < 0 || p1 > perLight) p1 = perLight;
  if (per_light) per_light->assign(s.lights.size(), 0);
  for (size_t li = 0; li < s.lights.size(); li++) {
    const Light& l = s.lights[li];
    const size_t before = list.size();
    for (int i = p0; i < p1; i++) {
      Rng g(seed, RTO_DOMAIN_PHOTON, (uint64_t)li * (uint64_t)perLight + (uint64_t)i);
      V3 startPosition = light_rand_area_position(l, g);
      V3 startDirection = hsphere_uniform_sample(g, l.normal);
      float pdf = dot(normalize(startDirection), normalize(l.normal));
      float weight = light_radiance(l, startPosition) / (pdf * lightPdf);
      Particle photon{mk(0.f, 0.f, 0.f), mk(0.f, 0.f, 0.f), weight};
      int depth = photon_path(s, startPosition, startDirection, photon, g, list, c);
      if (depth >= 0 && depth_hist) depth_hist[depth]++;
    }
    if (per_light) (*per_light)[li] = (int64_t)(list.size() - before);
  }
}

// Image.cpp:12-21
static inline V3 background(int y, int h) {
  float alpha = std::clamp((float)y / (float)(h - 1), 0.f, 1.f);
  V3 c0 = mk(0.1f, 0.2f, 0.8f), c1 = mk(0.9f, 0.9f, 1.0f);
  return add(scl(c0, 1.0f - alpha), scl(c1, alpha));  // Vec3.h:241-244 mix()
}
// Renderer.cpp:262-265 after the last sample
static inline V3 composite(V3 sum, int counter, V3 bg, int N) {
  return add(dvs(sum, (float)N), dvs(scl(bg, (float)(N - counter)), (float)N));
}

}  // namespace orc
#endif
