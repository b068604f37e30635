// scene_host.cpp -- see scene_host.h.  Strict IEEE (built with -ffp-contract=off): the camera, the light
// bases, the rotated vertices and the vertex normals feed bit-exact device geometry, so every expression
// keeps the operation order and the float/double promotions of the reference line it cites.
#include "scene_host.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <sys/stat.h>

namespace rth {
namespace {

inline Float3 operator+(Float3 a, Float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Float3 operator-(Float3 a, Float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Float3 operator*(Float3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(Float3 a, Float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // Vec3.h:220-223
inline Float3 cross(Float3 a, Float3 b) {                                            // Vec3.h:225-232
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline Float3 normalize(Float3 a) {  // Vec3.h:165-178: length through the double sqrt, then * (1/len)
  float len = (float)std::sqrt((double)dot(a, a));
  if (len == 0.0f) return a;
  float inv = 1.0f / len;
  return {a.x * inv, a.y * inv, a.z * inv};
}
inline void put(float* dst, Float3 v) {
  dst[0] = v.x;
  dst[1] = v.y;
  dst[2] = v.z;
}

// Mesh.h:126-134
void skip_hash_comment_line(std::ifstream& in) {
  while (in.peek() == '\n' || in.peek() == ' ') in.get();
  if (in.peek() == '#') {
    char trash[1024];
    in.getline(trash, 1023);
  }
}

rt_material material(float kd, float alpha, Float3 albedo, Float3 f0) {
  rt_material m;
  m.kd = kd;
  m.alpha = alpha;
  put(m.albedo, albedo);
  put(m.f0, f0);
  return m;
}

// createPlane (Main.cpp:26-37): 4 corners, the same normal 4 times, triangles (n, n+1, n+3), (n, n+2, n+3)
void add_plane(HostMesh& mesh, Float3 c0, Float3 c1, Float3 c2, Float3 c3, Float3 normal) {
  int n = (int)mesh.positions.size();
  const Float3 c[4] = {c0, c1, c2, c3};
  for (int i = 0; i < 4; i++) {
    mesh.positions.push_back(c[i]);
    mesh.normals.push_back(normal);
  }
  const int t[6] = {n, n + 1, n + 3, n, n + 2, n + 3};
  mesh.triangles.insert(mesh.triangles.end(), t, t + 6);
}

}  // namespace

void HostMesh::load_off(const std::string& filename) {
  try {
    positions.clear();
    triangles.clear();
    std::ifstream in(filename.c_str());
    if (!in) throw std::runtime_error("Error loading OFF file: " + filename);  // Mesh.h:62-64
    std::string tag;
    unsigned int num_v = 0, num_f = 0, num_e = 0;
    in >> tag;
    skip_hash_comment_line(in);
    in >> num_v >> num_f >> num_e;
    skip_hash_comment_line(in);
    positions.resize(num_v);
    for (unsigned int i = 0; i < num_v; i++) in >> positions[i].x >> positions[i].y >> positions[i].z;
    for (unsigned int f = 0; f < num_f; f++) {
      unsigned int s = 0;
      in >> s;
      std::vector<unsigned int> v(s);
      for (unsigned int j = 0; j < s; j++) in >> v[j];
      for (unsigned int j = 2; j < s; j++) {  // Mesh.h:80-81: fan around the first vertex
        triangles.push_back((int32_t)v[0]);
        triangles.push_back((int32_t)v[j - 1]);
        triangles.push_back((int32_t)v[j]);
      }
    }
  } catch (const std::exception& e) {
    throw std::runtime_error(std::string("Error Loading OFF file: ") + e.what());  // Mesh.h:84-88
  }
  recompute_normals();
}

bool HostMesh::load_off_cached(const std::string& filename, int subdivisions, const std::string& cache_dir) {
  auto plain = [&]() {
    load_off(filename);
    for (int i = 0; i < subdivisions; i++) subdivide();
  };
  struct stat st;
  if (cache_dir.empty() || stat(filename.c_str(), &st) != 0) {  // no cache, or let load_off report the missing file
    plain();
    return false;
  }
  std::string base = filename.substr(filename.find_last_of('/') == std::string::npos ? 0 : filename.find_last_of('/') + 1);
  std::ostringstream name;
  name << cache_dir << "/" << base << "." << (long long)st.st_size << "." << (long long)st.st_mtime << ".s" << subdivisions
       << ".offbin";
  const char magic[8] = {'O', 'F', 'F', 'B', 'I', 'N', '0', '1'};
  {
    std::ifstream in(name.str().c_str(), std::ios::binary);
    char m[8];
    uint64_t nv = 0, nt = 0;
    if (in && in.read(m, 8) && std::equal(m, m + 8, magic) && in.read((char*)&nv, 8) && in.read((char*)&nt, 8) &&
        nv < (1ull << 31) && nt < (1ull << 31)) {
      positions.resize(nv);
      normals.resize(nv);
      triangles.resize(3 * nt);
      in.read((char*)positions.data(), (std::streamsize)(nv * sizeof(Float3)));
      in.read((char*)normals.data(), (std::streamsize)(nv * sizeof(Float3)));
      in.read((char*)triangles.data(), (std::streamsize)(3 * nt * sizeof(int32_t)));
      if (in && in.peek() == std::ifstream::traits_type::eof()) return true;
    }
  }
  plain();
  mkdir(cache_dir.c_str(), 0755);  // best effort: a cache that cannot be written is only slower
  const std::string tmp = name.str() + ".tmp";
  std::ofstream out(tmp.c_str(), std::ios::binary);
  if (out) {
    const uint64_t nv = positions.size(), nt = triangles.size() / 3;
    out.write(magic, 8);
    out.write((const char*)&nv, 8);
    out.write((const char*)&nt, 8);
    out.write((const char*)positions.data(), (std::streamsize)(nv * sizeof(Float3)));
    out.write((const char*)normals.data(), (std::streamsize)(nv * sizeof(Float3)));
    out.write((const char*)triangles.data(), (std::streamsize)(3 * nt * sizeof(int32_t)));
    out.close();
    if (out)
      std::rename(tmp.c_str(), name.str().c_str());
    else
      std::remove(tmp.c_str());
  }
  return false;
}

void HostMesh::recompute_normals() {
  normals.resize(positions.size(), Float3{0.f, 0.f, 0.f});
  for (size_t t = 0; t + 2 < triangles.size(); t += 3) {
    const Float3 p0 = positions[triangles[t]], p1 = positions[triangles[t + 1]], p2 = positions[triangles[t + 2]];
    const Float3 nt = normalize(cross(p1 - p0, p2 - p0));  // Mesh.h:117-122
    for (int j = 0; j < 3; j++) normals[triangles[t + j]] = normals[triangles[t + j]] + nt;
  }
  for (Float3& n : normals) n = normalize(n);
}

void HostMesh::rotate_y(float phi) {
  const float c = std::cos(phi), s = std::sin(phi);  // float overloads, as `using namespace std` picks them
  const Float3 r0{c, 0.f, s}, r1{0.f, 1.f, 0.f}, r2{-s, 0.f, c};
  for (Float3& p : positions) p = Float3{dot(r0, p), dot(r1, p), dot(r2, p)};
}

void HostMesh::subdivide() {
  std::map<std::pair<int32_t, int32_t>, int32_t> mid;  // unique edge -> new vertex
  auto midpoint = [&](int32_t a, int32_t b) {
    std::pair<int32_t, int32_t> key(std::min(a, b), std::max(a, b));
    auto it = mid.find(key);
    if (it != mid.end()) return it->second;
    const Float3 pa = positions[key.first], pb = positions[key.second];
    positions.push_back(Float3{0.5f * (pa.x + pb.x), 0.5f * (pa.y + pb.y), 0.5f * (pa.z + pb.z)});
    int32_t idx = (int32_t)positions.size() - 1;
    mid.emplace(key, idx);
    return idx;
  };
  std::vector<int32_t> out;
  out.reserve(triangles.size() * 4);
  const size_t nt = triangles.size();
  for (size_t t = 0; t + 2 < nt; t += 3) {
    const int32_t a = triangles[t], b = triangles[t + 1], c = triangles[t + 2];
    const int32_t ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
    const int32_t tri[12] = {a, ab, ca, b, bc, ab, c, ca, bc, ab, bc, ca};
    out.insert(out.end(), tri, tri + 12);
  }
  triangles.swap(out);
  normals.clear();
  recompute_normals();
}

void HostScene::add_mesh(const HostMesh& m) {
  if (mesh_first_triangle.empty()) {
    mesh_first_triangle.push_back(0);
    mesh_first_vertex.push_back(0);
  }
  const int32_t vbase = (int32_t)(positions.size() / 3);
  for (size_t i = 0; i < m.positions.size(); i++) {
    const Float3 p = m.positions[i], n = i < m.normals.size() ? m.normals[i] : Float3{0.f, 0.f, 0.f};
    positions.insert(positions.end(), {p.x, p.y, p.z});
    normals.insert(normals.end(), {n.x, n.y, n.z});
  }
  for (int32_t v : m.triangles) triangles.push_back(v + vbase);
  materials.push_back(m.material);
  mesh_first_triangle.push_back((int32_t)(triangles.size() / 3));
  mesh_first_vertex.push_back((int32_t)(positions.size() / 3));
}

rt_scene HostScene::view() const {
  rt_scene s{};
  s.num_vertices = (int32_t)(positions.size() / 3);
  s.num_triangles = (int32_t)(triangles.size() / 3);
  s.num_meshes = (int32_t)materials.size();
  s.num_lights = (int32_t)lights.size();
  s.positions = positions.data();
  s.normals = normals.data();
  s.triangles = triangles.data();
  s.mesh_first_triangle = mesh_first_triangle.data();
  s.mesh_first_vertex = mesh_first_vertex.data();
  s.materials = materials.data();
  s.lights = lights.data();
  s.camera = camera;
  return s;
}

rt_camera make_camera(Float3 look_from, Float3 look_at, Float3 up, float vertical_fov_deg, float aspect) {
  const float angle = (float)((double)vertical_fov_deg * M_PI / (double)180.f);  // Camera.h:15
  const float half_height = (float)std::tan((double)(angle / 2.f));              // Camera.h:16 (::tan(double))
  const float half_width = aspect * half_height;
  const Float3 at_from = normalize(look_from - look_at);
  const Float3 u = normalize(cross(up, at_from));
  const Float3 v = cross(at_from, u);
  rt_camera c;
  put(c.position, look_from);
  put(c.lower_left, ((look_from - u * half_width) - v * half_height) - at_from);  // Camera.h:21
  put(c.horizontal, u * (2.f * half_width));                                      // Camera.h:22
  put(c.vertical, v * (2.f * half_height));                                       // Camera.h:23
  return c;
}

rt_light make_light(Float3 position, Float3 color, Float3 direction, float intensity, float side) {
  rt_light l;
  const Float3 normal = normalize(direction - position);                                               // LightSource.h:29
  const Float3 vertical = normalize(cross(normal, normalize(normal + Float3{1.f, 0.f, 0.f})));         // :30-31
  const Float3 horizontal = normalize(cross(normal, vertical));                                        // :32
  put(l.position, position);
  put(l.color, color);
  put(l.normal, normal);
  put(l.vertical, vertical);
  put(l.horizontal, horizontal);
  l.intensity = intensity;
  l.side = side;
  l.ac = 1.f;  // LightSource.h:65
  l.al = 0.3f;
  l.aq = 0.3f;
  l.factor = 4.5f;  // LightSource.h:63
  return l;
}

void build_reference_scene(int width, int height, const SceneOptions& opt, HostScene& out) {
  out = HostScene();
  // Main.cpp:169-170
  out.camera = make_camera({0.3f, 0.6f, 2.3f}, {0.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, 60.f,
                           (float)(size_t)width / (float)(size_t)height);
  // Main.cpp:101-124
  out.lights.push_back(make_light({-1.4f, 1.f, 2.9f}, {1.f, 1.f, 1.f}, {0.3f, 0.f, -1.f}, 0.85f, 0.01f));
  out.lights.push_back(make_light({1.4f, 1.f, 2.9f}, {1.f, 1.f, 1.f}, {-0.3f, 0.f, -1.f}, 0.85f, 0.01f));
  out.lights.push_back(make_light({0.f, -0.3f, 1.1f}, {1.f, 1.f, 1.f}, {0.f, 0.f, -1.f}, 0.85f, 0.1f));

  HostMesh walls, left_wall, right_wall, cube, cube2;
  // Main.cpp:126-151
  const Float3 walls_f0{0.5f, 0.5f, 0.5f};
  walls.material = material(0.6f, 0.3f, {0.96f, 0.96f, 0.86f}, walls_f0);
  left_wall.material = material(0.6f, 0.3f, {0.9f, 0.3f, 0.3f}, walls_f0);
  right_wall.material = material(0.6f, 0.3f, {0.3f, 0.9f, 0.3f}, walls_f0);
  cube.material = material(0.1f, 0.1f, {0.9f, 0.9f, 0.9f}, {1.0f, 0.86f, 0.57f});
  cube2.material = material(0.8f, 0.9f, {0.4f, 0.4f, 0.9f}, {(float)0.3, (float)0.3, (float)0.3});
  // Main.cpp:184-191; -i a.off,b.off,...: the list replaces mesh_cube, mesh_cube2 and appends further meshes
  std::vector<std::string> inputs;
  {
    std::stringstream list(opt.input_off);
    std::string item;
    while (std::getline(list, item, ','))
      if (!item.empty()) inputs.push_back(item);
  }
  cube.load_off_cached(inputs.size() > 0 ? inputs[0] : opt.mesh_dir + "/cube_tri.off", opt.subdivisions, opt.cache_dir);
  cube2.load_off_cached(inputs.size() > 1 ? inputs[1] : opt.mesh_dir + "/cube_tri2.off", 0, opt.cache_dir);
  std::vector<HostMesh> more(inputs.size() > 2 ? inputs.size() - 2 : 0);
  for (size_t i = 0; i < more.size(); i++) {
    more[i].material = (i % 2 == 0) ? cube.material : cube2.material;
    more[i].load_off_cached(inputs[i + 2], 0, opt.cache_dir);
  }
  // Main.cpp:39-86,194-196
  const float b = 1.51f, c = 1.5f;
  add_plane(walls, {b, -1.f, b}, {b, -1.f, -b}, {-b, -1.f, b}, {-b, -1.f, -b}, {0.f, 1.f, 0.f});    // ground
  add_plane(walls, {-b, -1.f, -b}, {b, -1.f, -b}, {-b, c, -b}, {b, c, -b}, {0.f, 0.f, 1.f});        // back wall
  add_plane(walls, {b, c, b}, {b, c, -b}, {-b, c, b}, {-b, c, -b}, {0.f, -1.f, 0.f});               // ceiling
  add_plane(left_wall, {-b, -1.f, b}, {-b, -1.f, -b}, {-b, c, b}, {-b, c, -b}, {1.f, 0.f, 0.f});
  add_plane(right_wall, {b, -1.f, b}, {b, -1.f, -b}, {b, c, b}, {b, c, -b}, {-1.f, 0.f, 0.f});
  // Main.cpp:201-202: float(M_PI / 4.5f)
  cube.rotate_y((float)(M_PI / 4.5f));
  cube2.rotate_y((float)(-M_PI / 4.5f));
  // Main.cpp:204-208: the mesh order is the tie-break order of RayTracer::rayTrace
  out.add_mesh(walls);
  out.add_mesh(left_wall);
  out.add_mesh(right_wall);
  out.add_mesh(cube);
  out.add_mesh(cube2);
  for (size_t i = 0; i < more.size(); i++) {
    more[i].rotate_y((float)((i % 2 == 0 ? 1.0 : -1.0) * (M_PI / 4.5f)));
    out.add_mesh(more[i]);
  }
}

void fill_background(int width, int height, std::vector<float>& rgb) {
  rgb.resize((size_t)width * height * 3);
  const float c0[3] = {0.1f, 0.2f, 0.8f}, c1[3] = {0.9f, 0.9f, 1.0f};
  for (int y = 0; y < height; y++) {
    const float alpha = std::clamp((float)y / (float)(size_t)(height - 1), 0.f, 1.f);  // Image.cpp:17-18
    for (int x = 0; x < width; x++)
      for (int ch = 0; ch < 3; ch++)
        rgb[3 * ((size_t)y * width + x) + ch] = c0[ch] * (1.0f - alpha) + c1[ch] * alpha;  // Vec3.h:241-244
  }
}

void save_ppm(const std::string& filename, int width, int height, const std::vector<float>& rgb) {
  std::ofstream out(filename.c_str());
  if (!out) {
    std::cerr << "Cannot open file " << filename.c_str() << std::endl;  // Image.cpp:25-28
    std::exit(1);
  }
  out << "P3" << std::endl << width << " " << height << std::endl << "255" << std::endl;
  for (size_t i = 0; i < (size_t)width * height * 3; i++) out << static_cast<unsigned int>(255.f * rgb[i]) << " ";
  out << std::endl;
  out.close();
}

void save_ppm_binary(const std::string& filename, int width, int height, const std::vector<float>& rgb) {
  std::ofstream out(filename.c_str(), std::ios::binary);
  if (!out) {
    std::cerr << "Cannot open file " << filename.c_str() << std::endl;
    std::exit(1);
  }
  out << "P6\n" << width << " " << height << "\n255\n";
  std::vector<unsigned char> bytes((size_t)width * height * 3);
  for (size_t i = 0; i < bytes.size(); i++) bytes[i] = (unsigned char)static_cast<unsigned int>(255.f * rgb[i]);
  out.write(reinterpret_cast<const char*>(bytes.data()), (std::streamsize)bytes.size());
  out.close();
}

void save_pcd(const std::string& filename, const std::vector<rt_photon>& photons) {
  std::ofstream out(filename.c_str());
  if (!out) {
    std::cerr << "Cannot open file " << filename.c_str() << std::endl;  // PhotonMap.h:61-64
    std::exit(1);
  }
  out << "VERSION .7" << std::endl
      << "FIELDS x y z normal_x normal_y normal_z" << std::endl
      << "SIZE 4 4 4 4 4 4" << std::endl
      << "TYPE F F F F F F" << std::endl
      << "COUNT 1 1 1 1 1 1" << std::endl
      << "WIDTH " << photons.size() << std::endl
      << "HEIGHT 1" << std::endl
      << "VIEWPOINT 0 0 0 1 0 0 0" << std::endl
      << "POINTS " << photons.size() << std::endl
      << "DATA ascii" << std::endl;
  for (const rt_photon& p : photons)
    out << p.position[0] << " " << p.position[1] << " " << p.position[2] << " " << p.direction[0] << " "
        << p.direction[1] << " " << p.direction[2] << " " << std::endl;
  std::cout << "Particle map was saved to: " << filename << std::endl;
  out.close();
}

}  // namespace rth
