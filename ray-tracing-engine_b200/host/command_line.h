// command_line.h -- the reference's command line (source/CommandLine.h:9-102): same flags, defaults,
// banner and error texts.  Additive options (absent = stock behaviour): -i/-input <a.off>[,<b.off>[,...]] replaces
// ../meshes/cube_tri.off (and cube_tri2.off, then appends further meshes), -meshdir <dir>, -cache <dir> (binary OFF
// cache), -subdiv <n>, -seed <n>, -device <n>, -brute, -update <n> (rewrite
// update.ppm every n sample passes as the reference does after every pass, Renderer.cpp:268-269; 0 = at the end),
// -p6 1 (binary P6 output instead of ASCII P3, same quantisation).
#pragma once
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

struct CommandLine {
  size_t width = 380, height = 270, numRays = 16, mode = 0, numPhotons = 0, k = 5;  // CommandLine.h:11-18
  std::string outputFilename = "output.ppm";
  // additive
  std::string input, meshDir = "../meshes", cacheDir;  // -cache <dir>: binary OFF cache (off by default)
  int subdiv = 0, device = 0, update = 0;
  bool knnExact = false;  // -knn exact: canonical exact k nearest photons (RT_FLAG_KNN_EXACT) instead of the
                          // reference's kdtree::knearest (default, "-knn reference")
  bool p6 = false;  // -p6 1: write binary P6 instead of the reference's ASCII P3
  unsigned long long seed = 1;
  bool brute = false;

  void printUsage(const char* command) const {
    std::cerr << "USAGE: " << command
              << " [-w/-width <image width>][-h/-height <image height>][-o/-output "
                 "<outputfilename>][-N/-n/-numRays <number of rays per "
                 "pixel>][-m/-mode <mode (0 for Ray tracing, 1 for Path "
                 "tracing)>][-p/-numPhotons <number of photons for a photon map. If "
                 "defined, photon map-based rendering is used.>][-k <number of "
                 "neighbours in photon mapping. Use only with -p/-numPhotons>]"
                 "[-i/-input <mesh.off>[,<mesh2.off>...]][-meshdir <dir>][-cache <dir>][-subdiv <n>][-seed <n>][-device <n>][-brute 1][-update <n>][-p6 1][-knn reference|exact]"
              << std::endl;
  }

  void parse(int argc, char** argv) {
    for (int i = 1; i < argc; i++) {
      const std::string a = argv[i];
      if (i == argc - 1) {  // CommandLine.h:50-57: a trailing flag is -help or an error
        if (a == "-help") {
          printUsage(argv[0]);
          std::exit(0);
        }
        throw std::runtime_error("Missing argument");
      }
      if (a == "-w" || a == "-width") width = std::atoi(argv[++i]);
      else if (a == "-h" || a == "-height") height = std::atoi(argv[++i]);
      else if (a == "-o" || a == "-output") outputFilename = argv[++i];
      else if (a == "-N" || a == "-n" || a == "-numRays") numRays = std::atoi(argv[++i]);
      else if (a == "-m" || a == "-mode") mode = std::atoi(argv[++i]);
      else if (a == "-p" || a == "-numPhotons") numPhotons = std::atoi(argv[++i]);
      else if (a == "-k") k = std::atoi(argv[++i]);
      else if (a == "-i" || a == "-input") input = argv[++i];
      else if (a == "-meshdir") meshDir = argv[++i];
      else if (a == "-cache") cacheDir = argv[++i];
      else if (a == "-subdiv") subdiv = std::atoi(argv[++i]);
      else if (a == "-seed") seed = std::strtoull(argv[++i], nullptr, 10);
      else if (a == "-device") device = std::atoi(argv[++i]);
      else if (a == "-brute") brute = std::atoi(argv[++i]) != 0;
      else if (a == "-update") update = std::atoi(argv[++i]);
      else if (a == "-p6") p6 = std::atoi(argv[++i]) != 0;
      else if (a == "-knn") knnExact = std::string(argv[++i]) == "exact";
      else throw std::runtime_error("Unknown argument <" + a + ">");
    }
    // CommandLine.h:78-96
    std::cout << "#########################" << std::endl << "Mode: ";
    if (mode == 1) {
      std::cout << "Path tracing" << std::endl;
    } else {
      mode = 0;
      std::cout << "Ray tracing" << std::endl;
    }
    std::cout << "Photon map ";
    if (numPhotons == 0)
      std::cout << "OFF" << std::endl;
    else
      std::cout << "ON with " << numPhotons << " photons. Number of searched neighbours equals " << k << std::endl;
    std::cout << "width: " << width << ", height: " << height << std::endl;
    std::cout << "Output image filename: " << outputFilename << std::endl;
    std::cout << "#########################" << std::endl << std::endl;
  }
};
