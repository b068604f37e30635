// main.cpp -- `RayTracer`, the reference's program (source/Main.cpp:153-235) with Renderer::render
// replaced by the CUDA path behind the C ABI.  Everything else happens on the host as before: command
// line, scene assembly from ../meshes/*.off, background fill, PPM output, the wall-clock line.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <vector>

#include "../../include/rt_b200.h"
#include "command_line.h"
#include "scene_host.h"

namespace {
[[noreturn]] void die(int rc) {
  // the reference lets std::logic_error("k is greater than ...") escape main (abort); we report and exit(1)
  std::cerr << rt_last_error() << " (rt_b200 status " << rc << ")" << std::endl;
  std::exit(1);
}
void printProgressBar(float prop) {  // Renderer.cpp:287-295
  int progress = (int)(50.0f * prop + 0.5f);
  std::string bar;
  for (int i = 0; i < progress; i++) bar += "█";
  std::cout << "Raytracing... [" << bar << std::string(50 - progress, ' ') << "] " << progress * 2 << "%\r" << std::flush;
}
}  // namespace

int main(int argc, char** argv) {
  CommandLine args;
  try {
    args.parse(argc, argv);
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    args.printUsage(argv[0]);
    std::exit(1);
  }
  auto begin = std::chrono::steady_clock::now();

  rth::HostScene scene;
  try {
    rth::SceneOptions opt;
    opt.mesh_dir = args.meshDir;
    opt.input_off = args.input;
    opt.subdivisions = args.subdiv;
    opt.cache_dir = args.cacheDir;
    rth::build_reference_scene((int)args.width, (int)args.height, opt, scene);
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;  // Main.cpp:188-191
    std::exit(1);
  }

  rt_params p{};
  p.width = (int32_t)args.width;
  p.height = (int32_t)args.height;
  p.num_rays = (int32_t)args.numRays;
  p.mode = (int32_t)args.mode;
  p.num_photons = (int32_t)args.numPhotons;
  p.k = (int32_t)args.k;
  p.seed = args.seed;
  p.flags = (args.brute ? RT_FLAG_BRUTE_FORCE : 0) | (args.knnExact ? RT_FLAG_KNN_EXACT : 0);

  rt_scene view = scene.view();
  rt_ctx* ctx = nullptr;
  int rc = rt_create(&view, &p, args.device, &ctx);
  if (rc) die(rc);

  std::vector<float> image;
  rth::fill_background((int)args.width, (int)args.height, image);  // Main.cpp:221

  if (args.numPhotons > 0) {  // Renderer.cpp:209-213 (+ PhotonMap.h:16-22 banner)
    std::cout << "Constructing a photon map with " << args.numPhotons << " photons" << std::endl;
    int32_t per_light = 0;
    rt_photons_per_light(ctx, &per_light);
    std::cout << "Emitting " << per_light << " photons per light source" << std::endl;
    std::vector<rt_photon> list((size_t)std::max(1, per_light * view.num_lights));
    std::vector<int64_t> counts((size_t)std::max(1, view.num_lights));
    int32_t hist[20];
    if ((rc = rt_emit_photons(ctx, 0, -1, list.data(), (int64_t)list.size(), counts.data(), hist))) die(rc);
    int64_t n = 0;
    for (int l = 0; l < view.num_lights; l++) n += counts[l];
    for (int i = 0; i < 20; i++)
      if (hist[i] > 0) std::cout << hist[i] << " photons with depth " << i << std::endl;  // PhotonMap.h:46-48
    std::cout << "Constructing a kd-tree for the photon map." << std::endl;
    if ((rc = rt_set_photons(ctx, list.data(), n))) die(rc);
    list.resize((size_t)n);
    // Main.cpp:216 writes pointcloud.pcd before rendering (header-only in the reference: it dumps an
    // empty member map); we write the real map -- a superset of that behaviour.
    rth::save_pcd("pointcloud.pcd", list);
  }

  printProgressBar(0.f);
  // Renderer.cpp:262-269 rewrites update.ppm (and the bar) after every sample pass; -update n does it every n
  // passes from a device-side composite of the samples so far, 0 only once at the end (same final file).
  struct Preview {
    int w, h;
    bool p6;
  } preview{(int)args.width, (int)args.height, args.p6};
  auto on_update = [](void* user, int32_t done, int32_t total, const float* rgb) {
    const Preview* pv = static_cast<const Preview*>(user);
    std::vector<float> snap(rgb, rgb + 3 * (size_t)pv->w * pv->h);
    (pv->p6 ? rth::save_ppm_binary : rth::save_ppm)("update.ppm", pv->w, pv->h, snap);
    printProgressBar(total > 0 ? (float)done / (float)total : 1.f);
  };
  if ((rc = rt_render_progressive(ctx, image.data(), args.update > 0 ? args.update : (int)args.numRays + 1, on_update,
                                  &preview)))
    die(rc);  // Main.cpp:224
  printProgressBar(1.f);
  std::cout << std::endl;
  (args.p6 ? rth::save_ppm_binary : rth::save_ppm)(args.outputFilename, (int)args.width, (int)args.height, image);  // Main.cpp:227

  rt_stats st{};
  rt_get_stats(ctx, &st);
  auto end = std::chrono::steady_clock::now();
  std::cout << "Total time is " << std::chrono::duration_cast<std::chrono::seconds>(end - begin).count() << "[s]"
            << std::endl;  // Main.cpp:228-232
  std::printf("B200: %llu rays in %.3f ms of kernels = %.1f Mrays/s (%llu kernel launches, BVH %d nodes)\n",
              (unsigned long long)st.rays, st.device_ms, st.device_ms > 0 ? st.rays / st.device_ms / 1e3 : 0.0,
              (unsigned long long)st.kernel_launches, st.bvh_nodes);
  rt_destroy(ctx);
  return 0;
}
