// scene_host.h -- host side of the application the render path lives in: mesh loading, scene assembly,
// image I/O.  These stay on the CPU in the B200 build exactly as in the reference (BASELINE.json
// north_star: "the C++ host code keeps the reference's RayTracer CLI, its .off mesh loading and its
// PPM output").  Behaviour (values, op order, error texts, file formats) follows the cited reference
// lines; the code is written from scratch against flat arrays because its only consumer is the C ABI.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"

namespace rth {

struct Float3 {
  float x, y, z;
};

// A triangle mesh as Mesh holds it (source/Mesh.h:136-139): positions, per-vertex normals, triangles
// with mesh-local vertex indices, one material.
struct HostMesh {
  std::vector<Float3> positions, normals;
  std::vector<int32_t> triangles;  // 3 per triangle
  rt_material material{};

  // Mesh::loadOFF (source/Mesh.h:57-90): polygons are fan-triangulated, '#' comment lines are skipped
  // where the reference skips them, normals recomputed.  Throws std::runtime_error with the
  // reference's messages.
  void load_off(const std::string& filename);
  // load_off + `subdivisions` midpoint subdivisions through a binary cache: `<cache_dir>/<name>.<size>.<mtime>.s<n>.offbin`
  // holds the loaded mesh (positions, normals, triangles as raw binary32 / int32 -- exactly what load_off and subdivide
  // produce, so a cached load is bit-identical) and is (re)written whenever it is missing or the .off file changed.
  // An empty cache_dir means no cache.  Parsing 23 MB of ASCII and subdividing to 1.2 M triangles takes seconds; the
  // cached load is one read.  Returns true when the mesh came from the cache.
  bool load_off_cached(const std::string& filename, int subdivisions, const std::string& cache_dir);
  // Mesh::recomputeNormals (source/Mesh.h:45-55): sum of UNIT face normals per vertex, normalised.
  void recompute_normals();
  // rotationY (source/Main.cpp:88-99): positions only -- the reference leaves the normals unrotated.
  void rotate_y(float phi);
  // one level of midpoint subdivision (1 triangle -> 4, one new vertex per unique edge at
  // 0.5f*(a+b)); used to synthesise the >= 1 M-triangle scene of BASELINE config 5.
  void subdivide();
};

// Scene (source/Scene.h) in the flat form the C ABI takes.  Owns its arrays; view() points into them.
struct HostScene {
  std::vector<float> positions, normals;
  std::vector<int32_t> triangles, mesh_first_triangle, mesh_first_vertex;
  std::vector<rt_material> materials;
  std::vector<rt_light> lights;
  rt_camera camera{};
  void add_mesh(const HostMesh& m);
  rt_scene view() const;
};

// Camera::Camera (source/Camera.h:9-24)
rt_camera make_camera(Float3 look_from, Float3 look_at, Float3 up, float vertical_fov_deg, float aspect);
// LightSource::LightSource (source/LightSource.h:19-33): basis from normalize(direction - position)
rt_light make_light(Float3 position, Float3 color, Float3 direction, float intensity, float side);

struct SceneOptions {
  std::string mesh_dir = "../meshes";  // the reference resolves ../meshes/ from its cwd (Main.cpp:186-187)
  // -i a.off[,b.off[,c.off...]]: the first file replaces cube_tri.off as mesh_cube, the second cube_tri2.off as
  // mesh_cube2; further files are appended after them as additional meshes, alternating the two cubes' materials and
  // rotations (Main.cpp:139-144,201-202) in scene order -- mesh order is the tie-break order of rayTrace
  std::string input_off;
  int subdivisions = 0;                // midpoint subdivisions applied to the FIRST input mesh
  std::string cache_dir;               // binary OFF cache directory ("" = off); see HostMesh::load_off_cached
};
// The scene main() assembles (source/Main.cpp:165-208): camera, 3 lights, Cornell box, two meshes.
void build_reference_scene(int width, int height, const SceneOptions& opt, HostScene& out);

// Image (source/Image.h, Image.cpp)
void fill_background(int width, int height, std::vector<float>& rgb);                      // Image.cpp:12-21
void save_ppm(const std::string& filename, int width, int height, const std::vector<float>& rgb);  // Image.cpp:23-43
// the same pixels (same `unsigned(255.f * v)` quantisation) as binary P6: 6 MB instead of ~23 MB of text at 1080p
void save_ppm_binary(const std::string& filename, int width, int height, const std::vector<float>& rgb);
// PhotonMap::saveToPCD (source/PhotonMap.h:59-84)
void save_pcd(const std::string& filename, const std::vector<rt_photon>& photons);

}  // namespace rth
