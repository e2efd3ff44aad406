// tests/host_harness.cu — CPU-side check of the engine's exact BVH search LOGIC (bvh.cuh compiled as host code: the AABB
// bound, the strict pruning rule, the (d2, index) ordering) against the oracle's kd-tree.  Test infrastructure: built by
// tests/test_host_logic.py with nvcc, never shipped.  The structure is built serially on the host with the same rules as
// k_morton_keys / k_bvh_leaves; the traversal is bvh_search_one, whose per-query decisions are exactly the per-lane
// decisions of the warp-group traversal used on the device.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>
#include "../hdl_graph_slam_b200/csrc/common.cuh"
#include "../hdl_graph_slam_b200/csrc/bvh.cuh"
#include "../oracle/oracle.h"

using namespace b2r;

#include "host_bvh.hpp"

struct HostKnn {
  static constexpr int kTileUnroll = 1;
  static constexpr int kTileLanes = 3;
  std::vector<float> d;
  std::vector<int> id;
  int k, cnt = 0;
  float worst() const { return cnt < k ? INFINITY : d[k - 1]; }
  float limit() const { return INFINITY; }
  void visit(float d2, int idx, int) {
    auto less = [&](int slot) { return d2 < d[slot] || (d2 == d[slot] && idx < id[slot]); };
    if (cnt == k && !less(k - 1)) return;
    int j = cnt < k ? cnt++ : k - 1;
    while (j > 0 && less(j - 1)) { d[j] = d[j - 1]; id[j] = id[j - 1]; j--; }
    d[j] = d2; id[j] = idx;
  }
};

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 20000;
  int nq = argc > 2 ? atoi(argv[2]) : 5000;
  int mode = argc > 3 ? atoi(argv[3]) : 0;  // 0 lidar-ish, 1 lattice (many exact ties), 2 clustered + far outliers
  unsigned long long s = 12345 + mode;
  std::vector<float> pts = make_points(n, mode, s), qs(nq * 4);
  for (int i = 0; i < nq; i++) {
    int j = (int)(urand(s) * n);
    double sc = (mode == 1) ? ((i & 1) ? 0.0 : 0.25) : 0.3;
    qs[i * 4] = pts[j * 4] + (float)((urand(s) - 0.5) * sc); qs[i * 4 + 1] = pts[j * 4 + 1] + (float)((urand(s) - 0.5) * sc);
    qs[i * 4 + 2] = pts[j * 4 + 2] + (float)((urand(s) - 0.5) * sc); qs[i * 4 + 3] = 1.f;
    if (i % 97 == 0) { qs[i * 4] += 400.f; }  // far outside the cloud
  }
  HostBvh H = build(pts, n);
  printf("bvh nleaf=%d nsup=%d\n", H.b.nleaf, H.b.nsup);
  const int k = 20;
  std::vector<int32_t> oidx((size_t)nq * k);
  std::vector<float> od2((size_t)nq * k);
  orc_knn(pts.data(), n, 4, qs.data(), nq, 4, k, oidx.data(), od2.data(), 0);
  long bad1 = 0, badk = 0, badlim = 0;
  for (int i = 0; i < nq; i++) {
    Nn1 v; v.reset(INFINITY);
    bvh_search_one(H.b, qs[i * 4], qs[i * 4 + 1], qs[i * 4 + 2], v);
    if (v.best_idx() != oidx[(size_t)i * k] || v.best_d2() != od2[(size_t)i * k]) {
      if (bad1 < 5) printf("1nn mismatch q%d: got (%g,%d) want (%g,%d)\n", i, v.best_d2(), v.best_idx(), od2[(size_t)i * k], oidx[(size_t)i * k]);
      bad1++;
    }
    // range-limited search (GICP: limit 6.25): result must agree whenever the true NN is inside the limit
    Nn1 w; w.reset(6.25f);
    bvh_search_one(H.b, qs[i * 4], qs[i * 4 + 1], qs[i * 4 + 2], w);
    bool want_valid = od2[(size_t)i * k] < 6.25f;
    bool got_valid = w.best_pos >= 0 && w.best_d2() < 6.25f;
    if (want_valid != got_valid || (want_valid && (w.best_idx() != oidx[(size_t)i * k] || w.best_d2() != od2[(size_t)i * k]))) badlim++;
    HostKnn K; K.k = k; K.d.resize(k); K.id.resize(k);
    bvh_search_one(H.b, qs[i * 4], qs[i * 4 + 1], qs[i * 4 + 2], K);
    for (int j = 0; j < k; j++)
      if (K.id[j] != oidx[(size_t)i * k + j] || K.d[j] != od2[(size_t)i * k + j]) { badk++; break; }
  }
  printf("queries=%d  1nn_mismatch=%ld  limited_mismatch=%ld  knn_mismatch=%ld\n", nq, bad1, badlim, badk);
  return (bad1 || badk || badlim) ? 1 : 0;
}
