// api.cpp -- C entry points of the CPU restatement (oracle/port/port.h).
//
// TEST INFRASTRUCTURE (oracle/): see the header of port.h.  The entry points mirror those of the
// reference harness (oracle/ref_harness.cpp, prefix ref_) so that one Python wrapper
// (oracle/oracle.py) drives both and tests can diff them call by call.
#include <cstring>
#include <thread>

#include "port.h"

using namespace orc;

namespace {
struct PortPhotonMap {
  std::vector<Particle> list;
  KdTree tree;
  int depth_hist[20] = {0};
  Counters counters;
  std::vector<int64_t> per_light;
};
inline V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
inline void st3(float* p, V3 v) {
  p[0] = v.x;
  p[1] = v.y;
  p[2] = v.z;
}
Counters g_counters;
}  // namespace

extern "C" {

// Flat scene exactly as it crosses the Renderer::render seam: lights carry their
// host-computed bases (21 floats), the camera its 4 vectors (12 floats).
void* orc_scene_from_flat(int V, int T, int M, int L, const float* pos, const float* nrm, const int32_t* tri,
                          const int32_t* mesh_tri_off, const int32_t* mesh_vtx_off, const float* mats,
                          const float* lights21, const float* cam12, int w, int h) {
  Scene* s = new Scene();
  s->V = V;
  s->T = T;
  s->M = M;
  s->L = L;
  s->w = w;
  s->h = h;
  for (int i = 0; i < V; i++) {
    s->P.push_back(ld3(pos + 3 * i));
    s->N.push_back(ld3(nrm + 3 * i));
  }
  s->tri.assign(tri, tri + 3 * (size_t)T);
  s->mesh_tri_off.assign(mesh_tri_off, mesh_tri_off + M + 1);
  s->mesh_vtx_off.assign(mesh_vtx_off, mesh_vtx_off + M + 1);
  s->tri_mesh.resize(T);
  for (int m = 0; m < M; m++)
    for (int t = mesh_tri_off[m]; t < mesh_tri_off[m + 1]; t++) s->tri_mesh[t] = m;
  for (int m = 0; m < M; m++) {
    const float* a = mats + 8 * m;
    s->mats.push_back(Material{a[0], a[1], ld3(a + 2), ld3(a + 5)});
  }
  for (int l = 0; l < L; l++) {
    const float* a = lights21 + 21 * l;
    s->lights.push_back(
        Light{ld3(a), ld3(a + 3), ld3(a + 6), ld3(a + 9), ld3(a + 12), a[15], a[16], a[17], a[18], a[19], a[20]});
  }
  s->cam = Camera{ld3(cam12), ld3(cam12 + 3), ld3(cam12 + 6), ld3(cam12 + 9)};
  return s;
}
void orc_scene_destroy(void* s) { delete static_cast<Scene*>(s); }

void orc_trace(void* sp, const float* rays, int64_t n, int32_t* hit, int32_t* mesh, int32_t* tri3, float* uvd,
               int32_t* tri_index) {
  const Scene& s = *static_cast<Scene*>(sp);
  for (int64_t i = 0; i < n; i++) {
    Hit h = ray_trace(s, ld3(rays + 6 * i), ld3(rays + 6 * i + 3));
    hit[i] = h.found;
    mesh[i] = h.found ? h.mesh : 0;
    for (int c = 0; c < 3; c++) tri3[3 * i + c] = h.found ? s.tri[3 * h.tri + c] - s.mesh_vtx_off[h.mesh] : 0;
    uvd[3 * i] = h.found ? h.u : 0.f;
    uvd[3 * i + 1] = h.found ? h.v : 0.f;
    uvd[3 * i + 2] = h.found ? h.d : 0.f;
    if (tri_index) tri_index[i] = h.found ? h.tri : -1;
  }
}

// Multi-threaded batch trace for the CPU baseline timing (rays are independent; the reference
// itself is single-threaded as built -- the thread count is reported with the number).
void orc_trace_mt(void* sp, const float* rays, int64_t n, int32_t* hit, int32_t* tri_index, float* uvd, int threads) {
  const Scene& s = *static_cast<Scene*>(sp);
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&, t]() {
      for (int64_t i = t; i < n; i += threads) {
        Hit h = ray_trace(s, ld3(rays + 6 * i), ld3(rays + 6 * i + 3));
        hit[i] = h.found;
        tri_index[i] = h.found ? h.tri : -1;
        uvd[3 * i] = h.found ? h.u : 0.f;
        uvd[3 * i + 1] = h.found ? h.v : 0.f;
        uvd[3 * i + 2] = h.found ? h.d : 0.f;
      }
    });
  for (auto& th : pool) th.join();
}

void orc_triangle_intersect(const float* in, int64_t n, int32_t* flag, float* uvt) {
  for (int64_t i = 0; i < n; i++) {
    const float* a = in + 15 * i;
    float u = 0, v = 0, t = 0;
    flag[i] = triangle_intersect(ld3(a), ld3(a + 3), ld3(a + 6), ld3(a + 9), ld3(a + 12), u, v, t);
    uvt[3 * i] = u;
    uvt[3 * i + 1] = v;
    uvt[3 * i + 2] = t;
  }
}

void orc_bsdf(const float* mat8, const float* in, int64_t n, float* out) {
  Material m{mat8[0], mat8[1], ld3(mat8 + 2), ld3(mat8 + 5)};
  for (int64_t i = 0; i < n; i++)
    st3(out + 3 * i, evaluate_color_response(m, ld3(in + 9 * i), ld3(in + 9 * i + 3), ld3(in + 9 * i + 6)));
}

void orc_light_eval(void* sp, int light, const float* pts, int64_t n, float* out) {
  const Scene& s = *static_cast<Scene*>(sp);
  for (int64_t i = 0; i < n; i++) st3(out + 3 * i, light_evaluate(s.lights[light], ld3(pts + 3 * i)));
}

void orc_light_sample(void* sp, int light, uint64_t seed, uint64_t domain, uint64_t index0, int64_t n, float* out) {
  const Scene& s = *static_cast<Scene*>(sp);
  for (int64_t i = 0; i < n; i++) {
    Rng g(seed, domain, index0 + i);
    st3(out + 3 * i, light_rand_area_position(s.lights[light], g));
  }
}

void orc_jitter(uint64_t seed, uint64_t domain, uint64_t index0, int64_t n, int sample, int nsamples, float* out2) {
  for (int64_t i = 0; i < n; i++) {
    Rng g(seed, domain, index0 + i);
    jitter_sample(g, sample, nsamples, out2[2 * i], out2[2 * i + 1]);
  }
}

void orc_hsphere(uint64_t seed, uint64_t domain, uint64_t index0, int64_t n, const float* normals, float* out3) {
  for (int64_t i = 0; i < n; i++) {
    Rng g(seed, domain, index0 + i);
    st3(out3 + 3 * i, hsphere_uniform_sample(g, ld3(normals + 3 * i)));
  }
}

void orc_rng_words(uint64_t seed, uint64_t domain, uint64_t index, int n, uint32_t* out) {
  Rng g(seed, domain, index);
  for (int i = 0; i < n; i++) out[i] = g.word();
}
void orc_rng_uniform_float(uint64_t seed, uint64_t domain, uint64_t index, int n, float a, float b, float* out) {
  Rng g(seed, domain, index);
  for (int i = 0; i < n; i++) out[i] = g.uniform_f(a, b);
}
void orc_rng_uniform_double(uint64_t seed, uint64_t domain, uint64_t index, int n, double a, double b, double* out) {
  Rng g(seed, domain, index);
  for (int i = 0; i < n; i++) out[i] = g.uniform_d(a, b);
}

void orc_camera_rays(void* sp, const int32_t* xy, const float* shift, int64_t n, float* rays) {
  const Scene& s = *static_cast<Scene*>(sp);
  for (int64_t i = 0; i < n; i++) {
    V3 o, d;
    camera_ray(s.cam, ((float)xy[2 * i] + shift[2 * i]) / (float)s.w,
               1.f - ((float)xy[2 * i + 1] + shift[2 * i + 1]) / (float)s.h, o, d);
    st3(rays + 6 * i, o);
    st3(rays + 6 * i + 3, d);
  }
}

// ---------------------------------------------------------------- photon map + kd-tree
void* orc_photon_map_create(void* sp, int numPhotons, uint64_t seed, int first_path, int num_paths) {
  const Scene& s = *static_cast<Scene*>(sp);
  PortPhotonMap* pm = new PortPhotonMap();
  int p0 = first_path < 0 ? 0 : first_path;
  int p1 = num_paths < 0 ? -1 : p0 + num_paths;
  emit_photons(s, numPhotons, seed, p0, p1, pm->list, pm->depth_hist, pm->counters, &pm->per_light);
  pm->tree.build(pm->list);
  return pm;
}
void* orc_photon_map_from_list(const float* particles, int64_t n) {
  PortPhotonMap* pm = new PortPhotonMap();
  for (int64_t i = 0; i < n; i++)
    pm->list.push_back(Particle{ld3(particles + 7 * i), ld3(particles + 7 * i + 3), particles[7 * i + 6]});
  pm->tree.build(pm->list);
  return pm;
}
void orc_photon_map_destroy(void* p) { delete static_cast<PortPhotonMap*>(p); }
int64_t orc_photon_map_size(void* p) { return (int64_t) static_cast<PortPhotonMap*>(p)->list.size(); }
// particles stored per light by the emission that built this map (out: one int64 per light)
void orc_photon_map_light_counts(void* p, int64_t* out, int n) {
  PortPhotonMap* pm = static_cast<PortPhotonMap*>(p);
  for (int i = 0; i < n; i++) out[i] = i < (int)pm->per_light.size() ? pm->per_light[i] : 0;
}
uint64_t orc_photon_map_rays(void* p) { return static_cast<PortPhotonMap*>(p)->counters.rays; }
void orc_photon_map_get(void* p, float* particles, int32_t* depth_hist) {
  PortPhotonMap* pm = static_cast<PortPhotonMap*>(p);
  for (size_t i = 0; i < pm->list.size(); i++) {
    st3(particles + 7 * i, pm->list[i].position);
    st3(particles + 7 * i + 3, pm->list[i].direction);
    particles[7 * i + 6] = pm->list[i].weight;
  }
  if (depth_hist)
    for (int i = 0; i < 20; i++) depth_hist[i] = pm->depth_hist[i];
}
void orc_kdtree_layout(void* p, float* nodes7, int32_t* left, int32_t* right, int32_t* root) {
  PortPhotonMap* pm = static_cast<PortPhotonMap*>(p);
  for (size_t i = 0; i < pm->tree.nodes.size(); i++) {
    const KdNode& n = pm->tree.nodes[i];
    st3(nodes7 + 7 * i, n.p.position);
    st3(nodes7 + 7 * i + 3, n.p.direction);
    nodes7[7 * i + 6] = n.p.weight;
    left[i] = n.left;
    right[i] = n.right;
  }
  *root = pm->tree.root;
}
// out7 as the reference harness; node_index (nullable) receives the kd-array index of each result.
int orc_knn(void* p, const float* q3, int64_t nq, int k, float* out7, int64_t* visited, int32_t* node_index) {
  PortPhotonMap* pm = static_cast<PortPhotonMap*>(p);
  if (pm->tree.empty() || k > (int)pm->tree.nodes.size()) return 1;  // kdtree.h:181-183 throws
  std::vector<int32_t> result;
  for (int64_t i = 0; i < nq; i++) {
    pm->tree.knearest(ld3(q3 + 3 * i), k, result);
    if (visited) visited[i] = (int64_t)pm->tree.visited;
    for (int j = 0; j < k; j++) {
      const Particle& a = pm->tree.nodes[result[j]].p;
      if (out7) {
        float* o = out7 + 7 * (i * k + j);
        st3(o, a.position);
        st3(o + 3, a.direction);
        o[6] = a.weight;
      }
      if (node_index) node_index[i * k + j] = result[j];
    }
  }
  return 0;
}

// ---------------------------------------------------------------- the render loop
// Same contract as ref_render (oracle/ref_harness.cpp).  threads > 1 splits the window rows over
// host threads (each (pixel, sample) owns its random stream, so the result does not depend on it).
int orc_render_mt(void* sp, int N, int mode, int numPhotons, int k, uint64_t seed, void* photon_map, int x0, int y0,
                  int x1, int y1, int s0, int s1, float* samples, int8_t* found, float* sum_rgb, int32_t* counter,
                  int threads) {
  const Scene& s = *static_cast<Scene*>(sp);
  PortPhotonMap* pm = static_cast<PortPhotonMap*>(photon_map);
  bool use_tree = pm && !pm->tree.empty();
  if (use_tree && k > (int)pm->tree.nodes.size()) return 1;
  int ww = x1 - x0, wh = y1 - y0;
  if (threads < 1) threads = 1;
  std::vector<Counters> cnt(threads);
  auto work = [&](int tid) {
    KdTree local;  // knearest mutates `visited`; give every thread its own view of the tree
    KdTree* tree = nullptr;
    if (use_tree) {
      if (threads == 1)
        tree = &pm->tree;
      else {
        local = pm->tree;
        tree = &local;
      }
    }
    for (int y = y0 + tid; y < y1; y += threads)
      for (int x = x0; x < x1; x++) {
        size_t wi = (size_t)(y - y0) * ww + (x - x0);
        for (int i = s0; i < s1; i++) {
          bool f = true;
          V3 c = render_sample(s, N, mode, tree, k, numPhotons, seed, x, y, i, f, cnt[tid]);
          if (f) counter[wi]++;
          sum_rgb[3 * wi] += c.x;
          sum_rgb[3 * wi + 1] += c.y;
          sum_rgb[3 * wi + 2] += c.z;
          if (samples) {
            size_t si = (size_t)(i - s0) * wh * ww + wi;
            st3(samples + 3 * si, c);
            if (found) found[si] = f;
          }
        }
      }
  };
  if (threads == 1)
    work(0);
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  for (const Counters& c : cnt) {
    g_counters.rays += c.rays;
    g_counters.queries += c.queries;
    g_counters.visits += c.visits;
  }
  return 0;
}
int orc_render(void* sp, int N, int mode, int numPhotons, int k, uint64_t seed, void* photon_map, int x0, int y0,
               int x1, int y1, int s0, int s1, float* samples, int8_t* found, float* sum_rgb, int32_t* counter) {
  return orc_render_mt(sp, N, mode, numPhotons, k, seed, photon_map, x0, y0, x1, y1, s0, s1, samples, found, sum_rgb,
                       counter, 1);
}
// rays / knn queries / kd visits accumulated by orc_render* since the last reset
void orc_counters(uint64_t out[3], int reset) {
  out[0] = g_counters.rays;
  out[1] = g_counters.queries;
  out[2] = g_counters.visits;
  if (reset) g_counters = Counters();
}

void orc_background(int w, int h, float* rgb) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) st3(rgb + 3 * ((size_t)y * w + x), background(y, h));
}
void orc_composite(int w, int h, int N, const float* sum_rgb, const int32_t* counter, float* bg_inout) {
  for (size_t p = 0; p < (size_t)w * h; p++)
    st3(bg_inout + 3 * p, composite(ld3(sum_rgb + 3 * p), counter[p], ld3(bg_inout + 3 * p), N));
}

}  // extern "C"
