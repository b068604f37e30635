// kd_build.h -- device builder of the canonical photon kd-tree (csrc/kd_build.cu), used by the exact k-NN mode.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace rtb {

// height of the median-split tree over n photons (kdtree.h:60-69: the left range gets floor(size/2) photons)
int kd_canonical_height(long long n);

// d_photons7: n particles of 7 floats in emission order (device).  Writes the kd-ordered node arrays
// (position + weight, incomeDirection) and, per array position, the particle's index in the emitted list.
// The tree orders the photons of a range by (coordinate, list index): identical to build_kdtree_canonical (host).
bool build_kdtree_device(const float* d_photons7, int n, float4* d_kd_pos, float4* d_kd_dir, int* d_kd_orig,
                         cudaStream_t st, int* height_out, long long* launches_out, std::string& err);

}  // namespace rtb
