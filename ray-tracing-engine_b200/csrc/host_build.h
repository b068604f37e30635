// host_build.h -- host-side builders that feed the device: the flattened BVH (reference split policy)
// and the photon kd-tree (reference procedure).  Plain C++ (no CUDA), compiled with strict IEEE flags.
#pragma once
#include <cstdint>
#include <vector>

namespace rtb {

// Flattened BVH.
//
// Topology follows the reference's BVH::from_triangles (source/BVH.h:100-161) per mesh -- node box =
// union of the triangles' vertices; leaf iff one triangle (:123); cut axis = first axis with the
// strictly largest extent (:131-140); triangles sorted by the sum of their three vertex coordinates on
// that axis (:141-150); left = first floor(n/2), right = the rest (:151-158) -- i.e. a full binary
// tree with one-triangle leaves.  The per-mesh trees (the reference builds one per Mesh,
// source/Mesh.h:107-114) are joined by a small top-level tree over the mesh boxes built with the same
// policy.  The reference's own traversal semantics are NOT reproduced:
// the device walks this tree with a conservative slab test and the (t, index) lexicographic
// acceptance, which returns exactly what the brute-force RayTracer::rayTrace returns.
//
// Layout: 16 floats (64 B, four 128-bit loads) per internal node
//   [0..2] child0.lo  [3..5] child0.hi  [6..8] child1.lo  [9..11] child1.hi
//   [12] child0 ref (int bits)  [13] child1 ref  [14],[15] unused
// ref >= 0: internal node index; ref < 0: leaf, triangle slot = ~ref.  Node 0 is the root; nodes are
// in depth-first pre-order.  Triangle slots are in leaf (depth-first) order.
//
// Padding: every child box is grown by pad = pad_fraction * extent on all sides, extent = largest
// |coordinate| of any vertex, light or camera position.  Moller-Trumbore in binary32 can accept a
// triangle whose exact intersection lies ~1e-6*extent outside it; the pad (default 2^-14 ~ 6e-5)
// is ~50x that bound and also dominates the rounding of the slab arithmetic (<= 2^-22 * extent *
// |1/d| against pad * |1/d|), so a triangle the brute force accepts is never culled.
struct Bvh {
  std::vector<float> nodes;        // 16 per node (device-built trees: only the top-level join, see bvh_build.h)
  size_t num_nodes = 0;            // nodes of the whole tree
  std::vector<int32_t> slot_tri;   // slot -> global triangle index
  int depth = 0;                   // longest root-to-leaf path in nodes (stack bound)
  float pad = 0.f;
  // The per-mesh roots in mesh order, 7 floats each: padded lo[3], hi[3], ref (int bits).  The reference keeps
  // one BVH per Mesh and loops over the meshes (source/RayTracer.h:56-85); a traversal may start from this list
  // instead of the top-level tree (6 root-box tests in a row instead of 6 dependent node visits).
  std::vector<float> roots;
  int mesh_depth = 0;              // longest root-to-leaf path inside one mesh's tree
};

void build_bvh(int num_vertices, const float* positions, int num_triangles, const int32_t* triangles, int num_meshes,
               const int32_t* mesh_first_triangle, float extent, float pad_fraction, Bvh& out);

// The top-level join alone, for a tree whose per-mesh subtrees were built elsewhere (csrc/bvh_build.cu): n_roots
// exact (unpadded) mesh boxes lo[3],hi[3] and their references, mesh_depth = nodes on the longest root-to-leaf path
// inside one mesh.  out.nodes must already hold >= max(n_roots-1, 1) nodes; the join is written to nodes
// [0, n_roots-1); out.roots / out.depth / out.mesh_depth / out.pad are filled.
void build_top_level(int n_roots, const float* boxes6, const int32_t* refs, float pad, int mesh_depth, Bvh& out);

// Photon kd-tree exactly as kdtree::make_tree builds it (source/kdtree.h:60-69,119-126): in-place,
// node at begin+(end-begin)/2 after std::nth_element on the cycling axis; the array ends in the
// tree's in-order layout and the links are implied by the ranges.  photons: 7 floats each.
void build_kdtree(std::vector<float>& photons7, int* height_out);
// The canonical tree of the exact k-NN mode: the same procedure with the photons of a range ordered by
// (coordinate, index in the emitted list) -- a total order, so the array is a function of the list alone and the device
// builder (csrc/kd_build.cu) produces the identical one.  orig_out[i] = list index of the photon at array position i.
void build_kdtree_canonical(std::vector<float>& photons7, int* height_out, std::vector<int32_t>* orig_out);
// explicit links for inspection (rt_get_kdtree)
void kdtree_links(int64_t n, std::vector<int32_t>& left, std::vector<int32_t>& right, int32_t* root);

}  // namespace rtb
