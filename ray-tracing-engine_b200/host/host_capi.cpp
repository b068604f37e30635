// host_capi.cpp -- C entry points of the host-side scene code (lib/librt_host.so, no CUDA) so that the
// Python mirror and the tests build exactly the scene the CLI builds.
#include <cstring>
#include <exception>
#include <string>

#include "scene_host.h"

namespace {
thread_local std::string g_err;
}

extern "C" {

const char* rth_last_error(void) { return g_err.c_str(); }

static void* scene_create(const char* mesh_dir, const char* input_off, int subdivisions, int width, int height,
                          const char* cache_dir) {
  rth::HostScene* s = new rth::HostScene();
  try {
    rth::SceneOptions opt;
    if (mesh_dir && mesh_dir[0]) opt.mesh_dir = mesh_dir;
    if (input_off) opt.input_off = input_off;
    if (cache_dir) opt.cache_dir = cache_dir;
    opt.subdivisions = subdivisions;
    rth::build_reference_scene(width, height, opt, *s);
  } catch (const std::exception& e) {
    g_err = e.what();
    delete s;
    return nullptr;
  }
  return s;
}
void* rth_scene_create(const char* mesh_dir, const char* input_off, int subdivisions, int width, int height) {
  return scene_create(mesh_dir, input_off, subdivisions, width, height, nullptr);
}
// the same with the binary OFF cache (SceneOptions::cache_dir); input_off may be a comma-separated list
void* rth_scene_create_cached(const char* mesh_dir, const char* input_off, int subdivisions, int width, int height,
                              const char* cache_dir) {
  return scene_create(mesh_dir, input_off, subdivisions, width, height, cache_dir);
}
void rth_scene_destroy(void* h) { delete static_cast<rth::HostScene*>(h); }
void rth_scene_counts(void* h, int32_t out[4]) {
  rt_scene v = static_cast<rth::HostScene*>(h)->view();
  out[0] = v.num_vertices;
  out[1] = v.num_triangles;
  out[2] = v.num_meshes;
  out[3] = v.num_lights;
}
void rth_scene_get(void* h, float* pos, float* nrm, int32_t* tri, int32_t* mesh_tri_off, int32_t* mesh_vtx_off,
                   float* mats8, float* lights21, float* cam12) {
  const rth::HostScene& s = *static_cast<rth::HostScene*>(h);
  std::memcpy(pos, s.positions.data(), s.positions.size() * 4);
  std::memcpy(nrm, s.normals.data(), s.normals.size() * 4);
  std::memcpy(tri, s.triangles.data(), s.triangles.size() * 4);
  std::memcpy(mesh_tri_off, s.mesh_first_triangle.data(), s.mesh_first_triangle.size() * 4);
  std::memcpy(mesh_vtx_off, s.mesh_first_vertex.data(), s.mesh_first_vertex.size() * 4);
  static_assert(sizeof(rt_material) == 32 && sizeof(rt_light) == 84 && sizeof(rt_camera) == 48, "POD layout");
  std::memcpy(mats8, s.materials.data(), s.materials.size() * sizeof(rt_material));
  std::memcpy(lights21, s.lights.data(), s.lights.size() * sizeof(rt_light));
  std::memcpy(cam12, &s.camera, sizeof(rt_camera));
}

// Mesh::loadOFF on a file: call with null arrays to size them.  0 ok, 1 = the loader threw (see rth_last_error)
int rth_load_off(const char* path, int subdivisions, int32_t counts[2], float* pos, float* nrm, int32_t* tri) {
  rth::HostMesh m;
  try {
    m.load_off(path);
    for (int i = 0; i < subdivisions; i++) m.subdivide();
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
  counts[0] = (int32_t)m.positions.size();
  counts[1] = (int32_t)(m.triangles.size() / 3);
  if (pos) std::memcpy(pos, m.positions.data(), m.positions.size() * sizeof(rth::Float3));
  if (nrm) std::memcpy(nrm, m.normals.data(), m.normals.size() * sizeof(rth::Float3));
  if (tri) std::memcpy(tri, m.triangles.data(), m.triangles.size() * 4);
  return 0;
}

void rth_camera(int width, int height, float* cam12) {
  rt_camera c = rth::make_camera({0.3f, 0.6f, 2.3f}, {0.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, 60.f,
                                 (float)(size_t)width / (float)(size_t)height);
  std::memcpy(cam12, &c, sizeof(c));
}
void rth_background(int width, int height, float* rgb) {
  std::vector<float> v;
  rth::fill_background(width, height, v);
  std::memcpy(rgb, v.data(), v.size() * 4);
}
void rth_save_ppm(const char* path, int width, int height, const float* rgb) {
  std::vector<float> v(rgb, rgb + (size_t)width * height * 3);
  rth::save_ppm(path, width, height, v);
}

// PhotonMap::saveToPCD (source/PhotonMap.h:59-84) for n particles of 7 floats
void rth_save_pcd(const char* path, const float* photons7, int64_t n) {
  std::vector<rt_photon> v((size_t)n);
  std::memcpy(v.data(), photons7, sizeof(rt_photon) * (size_t)n);
  rth::save_pcd(path, v);
}

void rth_save_ppm_binary(const char* path, int width, int height, const float* rgb) {
  std::vector<float> v(rgb, rgb + (size_t)width * height * 3);
  rth::save_ppm_binary(path, width, height, v);
}

}  // extern "C"
