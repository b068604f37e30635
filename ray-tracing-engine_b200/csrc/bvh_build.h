// bvh_build.h -- GPU build of the flattened BVH (reference split policy), see bvh_build.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "host_build.h"

namespace rtb {

// d_pos: float4 per vertex; d_vidx: (v0, v1, v2, mesh) per global triangle (both already on the device).
// Writes the per-mesh subtrees and the host-built top-level join into d_nodes (4 float4 per node, layout of
// host_build.h) and the leaf-order triangles into d_tris (3 float4 per slot).  `out` receives slot_tri, roots, depth,
// mesh_depth and pad; out.nodes is sized but holds only the join (the rest lives on the device).
// Returns false (with err set) when the scene is outside its range (T < 2, T >= 2^24, more than 256 meshes) or a
// CUDA call fails; the caller then falls back to build_bvh on the host.
bool build_bvh_device(const float4* d_pos, const int4* d_vidx, int T, int M, const int32_t* mesh_first_triangle,
                      float pad, float4* d_nodes, float4* d_tris, cudaStream_t st, Bvh& out, long long* launches_out,
                      std::string& err);

}  // namespace rtb
