/* ref_binding.h -- the reference-side binding of INTEGRATION.md section 2, as real code.
 *
 * TEST INFRASTRUCTURE (oracle/): proves the drop-in.  oracle/Makefile generates a patched copy of the reference's
 * Main.cpp into oracle/_ref/build/ (git-ignored; the reference's sources are never copied into the repository's
 * history) in which the ONE call `renderer.render(image);` (source/Main.cpp:224) becomes
 * `renderOnB200(scene, args, image);`, includes this header just before `int main`, and links the result against
 * lib/librt_b200.so.  Everything else in that program -- command line, OFF loading, scene assembly, background,
 * savePPM -- is the reference's own code, so its output.ppm must be byte-identical to what our from-scratch host
 * (bin/RayTracer) writes for the same arguments (tests/test_gpu_edges.py).
 *
 * Camera, LightSource and Material keep their members private (Camera.h:34-40, LightSource.h:61-65,
 * Material.h:62-64); the patched translation unit opens `private` the way oracle/ref_harness.cpp does.  A
 * maintainer would add three one-line accessors or a friend declaration instead.
 */
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "../include/rt_b200.h"

static rt_material toRtMaterial(const Material& m) {  // Material.h:62-64
  rt_material r;
  r.kd = m.m_kd;
  r.alpha = m.m_alpha;
  for (int c = 0; c < 3; c++) {
    r.albedo[c] = m.m_albedo[c];
    r.f0[c] = m.m_F0[c];
  }
  return r;
}
static rt_light toRtLight(const LightSource& l) {  // LightSource.h:61-65: the HOST-computed basis crosses the seam
  rt_light r;
  for (int c = 0; c < 3; c++) {
    r.position[c] = l.m_position[c];
    r.color[c] = l.m_color[c];
    r.normal[c] = l.m_normal[c];
    r.vertical[c] = l.m_vertical[c];
    r.horizontal[c] = l.m_horizontal[c];
  }
  r.intensity = l.m_intensity;
  r.side = l.m_sideLength;
  r.ac = l.ac;
  r.al = l.al;
  r.aq = l.aq;
  r.factor = l.m_factor;
  return r;
}
static rt_camera toRtCamera(const Camera& cam) {  // Camera.h:34-40
  rt_camera r;
  for (int c = 0; c < 3; c++) {
    r.position[c] = cam.m_position[c];
    r.lower_left[c] = cam.m_lowerLeftCorner[c];
    r.horizontal[c] = cam.m_horizontal[c];
    r.vertical[c] = cam.m_vertical[c];
  }
  return r;
}

// Flatten Scene (source/Scene.h) into the POD form the C ABI takes and run Renderer::render on the GPU.  Mesh and
// triangle order are kept: they are RayTracer::rayTrace's tie-break order (source/RayTracer.h:32-51).
static void renderOnB200(Scene& scene, const CommandLine& args, Image& image) {
  std::vector<float> pos, nrm;
  std::vector<int32_t> tri, triOff{0}, vtxOff{0};
  std::vector<rt_material> mats;
  std::vector<rt_light> lights;
  for (Mesh& m : scene.meshes()) {
    const int vbase = (int)pos.size() / 3;
    for (size_t i = 0; i < m.vertexPositions().size(); i++)
      for (int c = 0; c < 3; c++) {
        pos.push_back(m.vertexPositions()[i][c]);
        nrm.push_back(m.vertexNormals()[i][c]);
      }
    for (const Vec3i& t : m.indexedTriangles())
      for (int c = 0; c < 3; c++) tri.push_back(t[c] + vbase);
    triOff.push_back((int)tri.size() / 3);
    vtxOff.push_back((int)pos.size() / 3);
    mats.push_back(toRtMaterial(m.material()));
  }
  for (LightSource& l : scene.lightsources()) lights.push_back(toRtLight(l));
  rt_scene s{};
  s.num_vertices = (int32_t)pos.size() / 3;
  s.num_triangles = (int32_t)tri.size() / 3;
  s.num_meshes = (int32_t)mats.size();
  s.num_lights = (int32_t)lights.size();
  s.positions = pos.data();
  s.normals = nrm.data();
  s.triangles = tri.data();
  s.mesh_first_triangle = triOff.data();
  s.mesh_first_vertex = vtxOff.data();
  s.materials = mats.data();
  s.lights = lights.data();
  s.camera = toRtCamera(scene.camera());
  rt_params p{};
  p.width = (int32_t)args.width();
  p.height = (int32_t)args.height();
  p.num_rays = (int32_t)args.numRays();
  p.mode = (int32_t)args.mode();
  p.num_photons = (int32_t)args.numPhotons();
  p.k = (int32_t)args.k();
  p.seed = 1;
  rt_ctx* ctx = nullptr;
  if (rt_create(&s, &p, /*device*/ 0, &ctx) != RT_OK) {
    std::cerr << rt_last_error() << std::endl;
    exit(1);
  }
  // Image stores row-major Vec3f (source/Image.h:23-29) == W*H*3 floats, y = 0 on top
  static_assert(sizeof(Vec3f) == 3 * sizeof(float), "Vec3f must be three packed floats");
  if (rt_render(ctx, &image(0, 0)[0]) != RT_OK) {
    std::cerr << rt_last_error() << std::endl;
    exit(1);
  }
  rt_destroy(ctx);
}
