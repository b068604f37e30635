// host_build.cpp -- see host_build.h.  Host C++ only.
#include "host_build.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <utility>
#include <vector>

namespace rtb {
namespace {

struct Box {
  float lo[3], hi[3];
  void reset() {
    for (int a = 0; a < 3; a++) {
      lo[a] = FLT_MAX;
      hi[a] = -FLT_MAX;
    }
  }
  void grow(const Box& b) {
    for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], b.lo[a]);
      hi[a] = std::max(hi[a], b.hi[a]);
    }
  }
};

// BVH.h:131-140: first axis whose extent is strictly larger than everything before it
inline int cut_axis(const Box& b) {
  float longest = 0.f;
  int axis = 0;
  for (int a = 0; a < 3; a++) {
    float len = b.hi[a] - b.lo[a];
    if (len > longest) {
      longest = len;
      axis = a;
    }
  }
  return axis;
}

struct Builder {
  float pad;
  float* nodes;
  int32_t* slot_tri;
  std::vector<Box> tri_box;           // per global triangle, exact
  std::vector<float> key[3];          // BVH.h:143-148: p0[a] + p1[a] + p2[a]
  std::atomic<int> max_depth{0};

  void write_child(float* node, int side, const Box* b, int32_t ref) {
    float* lo = node + 6 * side;
    float* hi = lo + 3;
    for (int a = 0; a < 3; a++) {
      lo[a] = b ? b->lo[a] - pad : std::numeric_limits<float>::quiet_NaN();
      hi[a] = b ? b->hi[a] + pad : std::numeric_limits<float>::quiet_NaN();
    }
    std::memcpy(node + 12 + side, &ref, 4);
  }
  void note_depth(int d) {
    int cur = max_depth.load();
    while (d > cur && !max_depth.compare_exchange_weak(cur, d)) {
    }
  }

  // Subtree over the triangles ord[a][b..e) (the same set in each of the three arrays, sorted by key[a]):
  // internal nodes [node_base, node_base+n-1), slots [slot_base, slot_base+n).
  //
  // BVH.h:141-158 sorts the node's triangles by the key of the cut axis and gives the first floor(n/2) to the
  // left child.  Sorting every node from scratch is O(T log^2 T); here the three orders are produced once and
  // every split is a stable partition of the two other orders (O(n) per node, O(T log T) in total).  The order
  // among equal keys differs from libstdc++'s unstable std::sort -- it is unspecified in the reference as well,
  // and any choice gives a valid tree of the same shape (hit results do not depend on it).
  int32_t* ord[3] = {nullptr, nullptr, nullptr};
  int32_t* tmp = nullptr;        // partition scratch, same indexing as ord[]
  unsigned char* side = nullptr;  // per global triangle: 1 = goes right at the split being processed

  int32_t build(int b, int e, int node_base, int slot_base, int depth, Box& box, int fork_levels) {
    const int n = e - b;
    if (n == 1) {
      slot_tri[slot_base] = ord[0][b];
      box = tri_box[ord[0][b]];
      note_depth(depth);
      return ~slot_base;
    }
    box.reset();
    for (int i = b; i < e; i++) box.grow(tri_box[ord[0][i]]);
    const int axis = cut_axis(box);
    const int nl = n / 2;  // BVH.h:151-158
    if (n > 2) {           // n == 2: every order already has one triangle per side after the marks below
      for (int i = b; i < b + nl; i++) side[ord[axis][i]] = 0;
      for (int i = b + nl; i < e; i++) side[ord[axis][i]] = 1;
      for (int a = 0; a < 3; a++) {
        if (a == axis) continue;
        int32_t* o = ord[a];
        int l = b, r = b + nl;
        for (int i = b; i < e; i++) {
          const int32_t t = o[i];
          if (side[t]) tmp[r++] = t; else tmp[l++] = t;
        }
        std::memcpy(o + b, tmp + b, sizeof(int32_t) * (size_t)n);
      }
    } else {
      const int32_t t0 = ord[axis][b], t1 = ord[axis][b + 1];
      for (int a = 0; a < 3; a++) {
        ord[a][b] = t0;
        ord[a][b + 1] = t1;
      }
    }
    Box bl, br;
    int32_t rl, rr;
    if (fork_levels > 0 && n > 4096) {
      std::thread t([&]() { rl = build(b, b + nl, node_base + 1, slot_base, depth + 1, bl, fork_levels - 1); });
      rr = build(b + nl, e, node_base + nl, slot_base + nl, depth + 1, br, fork_levels - 1);
      t.join();
    } else {
      rl = build(b, b + nl, node_base + 1, slot_base, depth + 1, bl, 0);
      rr = build(b + nl, e, node_base + nl, slot_base + nl, depth + 1, br, 0);
    }
    float* node = nodes + 16 * (size_t)node_base;
    write_child(node, 0, &bl, rl);
    write_child(node, 1, &br, rr);
    node[14] = node[15] = 0.f;
    return node_base;
  }

  struct Item {
    int32_t ref;
    Box box;
    float key[3];
  };
  // top level over mesh subtrees: same policy on the mesh boxes (key = lo + hi on the cut axis)
  int32_t build_top(Item* items, int n, int node_base, int depth, Box& box, int sub_depth) {
    if (n == 1) {
      box = items[0].box;
      return items[0].ref;
    }
    box.reset();
    for (int i = 0; i < n; i++) box.grow(items[i].box);
    int axis = cut_axis(box);
    std::sort(items, items + n, [axis](const Item& a, const Item& b) { return a.key[axis] < b.key[axis]; });
    int nl = n / 2;
    Box bl, br;
    int32_t rl = build_top(items, nl, node_base + 1, depth + 1, bl, sub_depth);
    int32_t rr = build_top(items + nl, n - nl, node_base + nl, depth + 1, br, sub_depth);
    float* node = nodes + 16 * (size_t)node_base;
    write_child(node, 0, &bl, rl);
    write_child(node, 1, &br, rr);
    node[14] = node[15] = 0.f;
    note_depth(depth + 1 + sub_depth);
    return node_base;
  }
};

}  // namespace

void build_bvh(int num_vertices, const float* P, int T, const int32_t* tri, int M, const int32_t* mesh_first_triangle,
               float extent, float pad_fraction, Bvh& out) {
  (void)num_vertices;
  Builder B;
  out.pad = B.pad = extent * pad_fraction;
  B.tri_box.resize(T);
  for (int a = 0; a < 3; a++) B.key[a].resize(T);
  for (int t = 0; t < T; t++) {
    Box& b = B.tri_box[t];
    b.reset();
    const float* p[3] = {P + 3 * (size_t)tri[3 * t], P + 3 * (size_t)tri[3 * t + 1], P + 3 * (size_t)tri[3 * t + 2]};
    for (int a = 0; a < 3; a++) {
      for (int v = 0; v < 3; v++) {
        b.lo[a] = std::min(b.lo[a], p[v][a]);
        b.hi[a] = std::max(b.hi[a], p[v][a]);
      }
      B.key[a][t] = p[0][a] + p[1][a] + p[2][a];
    }
  }
  int nonempty = 0;
  for (int m = 0; m < M; m++)
    if (mesh_first_triangle[m + 1] > mesh_first_triangle[m]) nonempty++;
  size_t num_nodes = (size_t)std::max(T - nonempty, 0) + (size_t)std::max(nonempty - 1, 0);
  if (num_nodes == 0) num_nodes = 1;  // T <= 1: a root with at most one leaf child
  out.nodes.assign(16 * num_nodes, 0.f);
  out.num_nodes = num_nodes;
  out.slot_tri.assign(T, 0);
  B.nodes = out.nodes.data();
  B.slot_tri = out.slot_tri.data();

  // the three key orders, per mesh (a mesh's triangles are contiguous in [t0, t0+n))
  std::vector<int32_t> ord_store[3], tmp_store(T);
  std::vector<unsigned char> side_store(T, 0);
  unsigned hw = std::thread::hardware_concurrency();
  int fork_levels = hw >= 16 ? 4 : hw >= 8 ? 3 : hw >= 4 ? 2 : hw >= 2 ? 1 : 0;
  {
    auto sort_axis = [&](int a) {
      std::vector<std::pair<float, int32_t>> kv(T);
      for (int t = 0; t < T; t++) kv[t] = {B.key[a][t], t};
      for (int m = 0; m < M; m++) {
        int t0 = mesh_first_triangle[m], t1 = mesh_first_triangle[m + 1];
        if (t1 - t0 > 1)
          std::sort(kv.begin() + t0, kv.begin() + t1,
                    [](const std::pair<float, int32_t>& x, const std::pair<float, int32_t>& y) { return x.first < y.first; });
      }
      ord_store[a].resize(T);
      for (int t = 0; t < T; t++) ord_store[a][t] = kv[t].second;
    };
    if (T > 4096 && hw >= 3) {
      std::thread t1([&]() { sort_axis(1); }), t2([&]() { sort_axis(2); });
      sort_axis(0);
      t1.join();
      t2.join();
    } else {
      for (int a = 0; a < 3; a++) sort_axis(a);
    }
  }
  for (int a = 0; a < 3; a++) B.ord[a] = ord_store[a].data();
  B.tmp = tmp_store.data();
  B.side = side_store.data();
  std::vector<Builder::Item> items;
  int node_base = std::max(nonempty - 1, 0);
  int sub_depth = 0;
  for (int m = 0; m < M; m++) {
    int t0 = mesh_first_triangle[m], n = mesh_first_triangle[m + 1] - t0;
    if (n <= 0) continue;
    Builder::Item it;
    B.max_depth = 0;
    it.ref = B.build(t0, t0 + n, node_base, t0, 0, it.box, fork_levels);
    sub_depth = std::max(sub_depth, B.max_depth.load());
    for (int a = 0; a < 3; a++) it.key[a] = it.box.lo[a] + it.box.hi[a];
    items.push_back(it);
    node_base += n - 1;
  }
  out.mesh_depth = sub_depth + 1;
  out.roots.clear();
  for (const Builder::Item& it : items) {  // before build_top reorders them
    for (int a = 0; a < 3; a++) out.roots.push_back(it.box.lo[a] - B.pad);
    for (int a = 0; a < 3; a++) out.roots.push_back(it.box.hi[a] + B.pad);
    float ref_bits;
    std::memcpy(&ref_bits, &it.ref, 4);
    out.roots.push_back(ref_bits);
  }
  B.max_depth = sub_depth;
  if (items.empty()) {
    B.write_child(B.nodes, 0, nullptr, -1);
    B.write_child(B.nodes, 1, nullptr, -1);
  } else if (items.size() == 1 && items[0].ref < 0) {  // a single triangle in the whole scene
    B.write_child(B.nodes, 0, &items[0].box, items[0].ref);
    B.write_child(B.nodes, 1, nullptr, -1);
  } else if (items.size() > 1) {
    Box root;
    B.build_top(items.data(), (int)items.size(), 0, 0, root, sub_depth);
  }  // else: the single mesh's root already sits at node 0
  out.depth = B.max_depth.load() + 1;
}

namespace {
struct KdItem {
  float v[7];
};
// kdtree.h:60-69.  The two recursive calls work on disjoint ranges, so the upper levels fork threads: every
// std::nth_element call sees exactly the input it sees in the serial recursion, and the array ends up identical.
int make_tree(KdItem* nodes, size_t begin, size_t end, size_t index, int fork_levels) {
  if (end <= begin) return 0;
  size_t n = begin + (end - begin) / 2;
  std::nth_element(nodes + begin, nodes + n, nodes + end,
                   [index](const KdItem& a, const KdItem& b) { return a.v[index] < b.v[index]; });
  index = (index + 1) % 3;
  int hl = 0, hr = 0;
  if (fork_levels > 0 && end - begin > 16384) {
    std::thread t([&]() { hl = make_tree(nodes, begin, n, index, fork_levels - 1); });
    hr = make_tree(nodes, n + 1, end, index, fork_levels - 1);
    t.join();
  } else {
    hl = make_tree(nodes, begin, n, index, 0);
    hr = make_tree(nodes, n + 1, end, index, 0);
  }
  return 1 + std::max(hl, hr);
}
int32_t link_tree(size_t begin, size_t end, int32_t* left, int32_t* right) {
  if (end <= begin) return -1;
  size_t n = begin + (end - begin) / 2;
  left[n] = link_tree(begin, n, left, right);
  right[n] = link_tree(n + 1, end, left, right);
  return (int32_t)n;
}
}  // namespace

void build_top_level(int n_roots, const float* boxes6, const int32_t* refs, float pad, int mesh_depth, Bvh& out) {
  Builder B;
  B.pad = pad;
  out.pad = pad;
  B.nodes = out.nodes.data();
  std::vector<Builder::Item> items((size_t)n_roots);
  out.roots.clear();
  for (int r = 0; r < n_roots; r++) {
    Builder::Item& it = items[r];
    it.ref = refs[r];
    for (int a = 0; a < 3; a++) {
      it.box.lo[a] = boxes6[6 * r + a];
      it.box.hi[a] = boxes6[6 * r + 3 + a];
      it.key[a] = it.box.lo[a] + it.box.hi[a];
    }
    for (int a = 0; a < 3; a++) out.roots.push_back(it.box.lo[a] - pad);
    for (int a = 0; a < 3; a++) out.roots.push_back(it.box.hi[a] + pad);
    float ref_bits;
    std::memcpy(&ref_bits, &it.ref, 4);
    out.roots.push_back(ref_bits);
  }
  const int sub_depth = mesh_depth - 1;
  out.mesh_depth = mesh_depth;
  B.max_depth = sub_depth;
  if (n_roots > 1) {
    Box root;
    B.build_top(items.data(), n_roots, 0, 0, root, sub_depth);
  }
  out.depth = B.max_depth.load() + 1;
}

namespace {
struct KdItemIdx {
  float v[7];
  int32_t idx;
};
int make_tree_canonical(KdItemIdx* nodes, size_t begin, size_t end, size_t index, int fork_levels) {
  if (end <= begin) return 0;
  size_t n = begin + (end - begin) / 2;
  std::nth_element(nodes + begin, nodes + n, nodes + end, [index](const KdItemIdx& a, const KdItemIdx& b) {
    return a.v[index] < b.v[index] || (a.v[index] == b.v[index] && a.idx < b.idx);
  });
  index = (index + 1) % 3;
  int hl = 0, hr = 0;
  if (fork_levels > 0 && end - begin > 16384) {
    std::thread t([&]() { hl = make_tree_canonical(nodes, begin, n, index, fork_levels - 1); });
    hr = make_tree_canonical(nodes, n + 1, end, index, fork_levels - 1);
    t.join();
  } else {
    hl = make_tree_canonical(nodes, begin, n, index, 0);
    hr = make_tree_canonical(nodes, n + 1, end, index, 0);
  }
  return 1 + std::max(hl, hr);
}
}  // namespace

void build_kdtree_canonical(std::vector<float>& photons7, int* height_out, std::vector<int32_t>* orig_out) {
  const size_t n = photons7.size() / 7;
  std::vector<KdItemIdx> items(n);
  for (size_t i = 0; i < n; i++) {
    std::memcpy(items[i].v, &photons7[7 * i], 28);
    items[i].idx = (int32_t)i;
  }
  unsigned hw = std::thread::hardware_concurrency();
  int h = make_tree_canonical(items.data(), 0, n, 0, hw >= 16 ? 4 : hw >= 8 ? 3 : hw >= 4 ? 2 : hw >= 2 ? 1 : 0);
  if (orig_out) orig_out->resize(n);
  for (size_t i = 0; i < n; i++) {
    std::memcpy(&photons7[7 * i], items[i].v, 28);
    if (orig_out) (*orig_out)[i] = items[i].idx;
  }
  if (height_out) *height_out = h;
}

void build_kdtree(std::vector<float>& photons7, int* height_out) {
  static_assert(sizeof(KdItem) == 28, "Particle is 28 bytes");
  size_t n = photons7.size() / 7;
  unsigned hw = std::thread::hardware_concurrency();
  int h = make_tree(reinterpret_cast<KdItem*>(photons7.data()), 0, n, 0, hw >= 16 ? 4 : hw >= 8 ? 3 : hw >= 4 ? 2 : hw >= 2 ? 1 : 0);
  if (height_out) *height_out = h;
}

void kdtree_links(int64_t n, std::vector<int32_t>& left, std::vector<int32_t>& right, int32_t* root) {
  left.assign(n, -1);
  right.assign(n, -1);
  *root = link_tree(0, (size_t)n, left.data(), right.data());
}

}  // namespace rtb
