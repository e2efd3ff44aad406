// oracle/kdtree.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// Exact k-nearest-neighbour search with float32 distance semantics, standing in for
// pcl::search::KdTree / FLANN KDTreeSingleIndex<L2_Simple<float>> used by the reference's registration
// handle (reference call sites: apps/scan_matching_odometry_nodelet.cpp:316 nearestKSearch,
// src/hdl_graph_slam/information_matrix_calculator.cpp:50-66, fast_gicp update_correspondences /
// calculate_covariances [SURVEY.md A.4]).
//
// Semantics pinned by this oracle (SURVEY.md §7 "Bit-exact correspondences"):
//   d2(q,p) = ((qx-px)^2 + (qy-py)^2) + (qz-pz)^2   evaluated in float32, no FMA contraction
//   result  = the k lexicographically smallest (d2, index) pairs, ascending — ties go to the LOWEST index
// An exact search returns the same set whatever the tree shape, so a simple median-split tree is enough.
#pragma once
#include <vector>
#include <cstdint>
#include <cmath>
#include <algorithm>
#include <limits>

namespace orc {

struct KdTree {
  struct Node {
    float lo[3], hi[3];  // tight bounding box of the subtree's points
    int left, right;     // children (or -1)
    int begin, end;      // range in perm_ (leaves)
  };
  std::vector<Node> nodes;
  std::vector<int> perm;
  const float* pts = nullptr;  // base pointer of x; stride in floats
  size_t stride = 0;
  size_t n = 0;
  static constexpr int kLeaf = 12;

  inline const float* P(int i) const { return pts + (size_t)i * stride; }

  void build(const float* points, size_t n_, size_t stride_floats) {
    pts = points;
    stride = stride_floats;
    n = n_;
    // pcl::KdTreeFLANN::convertCloudToArray indexes the FINITE points only (non-dense clouds): a NaN / inf point is never a neighbour
    perm.clear();
    perm.reserve(n);
    for (size_t i = 0; i < n; i++) {
      const float* p = points + i * stride_floats;
      if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2])) perm.push_back((int)i);
    }
    nodes.clear();
    nodes.reserve(n / 4 + 16);
    if (!perm.empty()) build_rec(0, (int)perm.size());
  }

  int build_rec(int b, int e) {
    Node nd;
    for (int d = 0; d < 3; d++) {
      nd.lo[d] = std::numeric_limits<float>::infinity();
      nd.hi[d] = -std::numeric_limits<float>::infinity();
    }
    for (int i = b; i < e; i++) {
      const float* p = P(perm[i]);
      for (int d = 0; d < 3; d++) {
        nd.lo[d] = std::min(nd.lo[d], p[d]);
        nd.hi[d] = std::max(nd.hi[d], p[d]);
      }
    }
    nd.left = nd.right = -1;
    nd.begin = b;
    nd.end = e;
    int id = (int)nodes.size();
    nodes.push_back(nd);
    if (e - b > kLeaf) {
      int dim = 0;
      float ext = nd.hi[0] - nd.lo[0];
      for (int d = 1; d < 3; d++)
        if (nd.hi[d] - nd.lo[d] > ext) { ext = nd.hi[d] - nd.lo[d]; dim = d; }
      int mid = (b + e) / 2;
      std::nth_element(perm.begin() + b, perm.begin() + mid, perm.begin() + e, [&](int a, int c) {
        float va = P(a)[dim], vc = P(c)[dim];
        return va < vc || (va == vc && a < c);
      });
      int l = build_rec(b, mid);
      int r = build_rec(mid, e);
      nodes[id].left = l;
      nodes[id].right = r;
    }
    return id;
  }

  static inline float dist2(const float* q, const float* p) {
    float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
    float a = dx * dx, b = dy * dy, c = dz * dz;
    float s = a + b;
    return s + c;
  }
  // lower bound on dist2(q, any point in box): same evaluation order => monotone => never above a real d2
  static inline float box_dist2(const float* q, const Node& nd) {
    float b[3];
    for (int d = 0; d < 3; d++) {
      float v = 0.f;
      if (q[d] < nd.lo[d]) v = nd.lo[d] - q[d];
      else if (q[d] > nd.hi[d]) v = q[d] - nd.hi[d];
      b[d] = v;
    }
    float a = b[0] * b[0], bb = b[1] * b[1], c = b[2] * b[2];
    float s = a + bb;
    return s + c;
  }

  struct Heap {  // sorted ascending list of up to k (d2, idx)
    int k, cnt;
    float* d;
    int* id;
    inline bool better(float dd, int ii, int slot) const { return dd < d[slot] || (dd == d[slot] && ii < id[slot]); }
    inline float worst() const { return cnt < k ? std::numeric_limits<float>::infinity() : d[k - 1]; }
    inline void push(float dd, int ii) {
      if (cnt == k && !better(dd, ii, k - 1)) return;
      int pos = (cnt < k) ? cnt++ : k - 1;
      while (pos > 0 && better(dd, ii, pos - 1)) {
        d[pos] = d[pos - 1];
        id[pos] = id[pos - 1];
        pos--;
      }
      d[pos] = dd;
      id[pos] = ii;
    }
  };

  void search_rec(int node, const float* q, Heap& h) const {
    const Node& nd = nodes[node];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; i++) {
        int pi = perm[i];
        h.push(dist2(q, P(pi)), pi);
      }
      return;
    }
    float dl = box_dist2(q, nodes[nd.left]), dr = box_dist2(q, nodes[nd.right]);
    int first = nd.left, second = nd.right;
    float df = dl, ds = dr;
    if (dr < dl) { first = nd.right; second = nd.left; df = dr; ds = dl; }
    if (df <= h.worst()) search_rec(first, q, h);   // '<=' keeps equal-distance lower-index candidates reachable
    if (ds <= h.worst()) search_rec(second, q, h);
  }

  // returns number found (min(k, n)); out_idx/out_d2 ascending by (d2, idx)
  int knn(const float* q, int k, int* out_idx, float* out_d2) const {
    Heap h{k, 0, out_d2, out_idx};
    if (nodes.empty()) return 0;
    if (!(std::isfinite(q[0]) && std::isfinite(q[1]) && std::isfinite(q[2]))) return 0;  // PCL: a non-finite query finds nothing
    search_rec(0, q, h);
    return h.cnt;
  }
};

}  // namespace orc
