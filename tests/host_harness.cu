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

struct HostBvh {
  std::vector<float4> sp, llo, lhi, slo, shi;
  Bvh b;
};

static HostBvh build(const std::vector<float>& pts, int n) {
  HostBvh H;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], pts[i * 4 + d]); mx[d] = std::max(mx[d], pts[i * 4 + d]); }
  float ext = std::max(std::max(mx[0] - mn[0], mx[1] - mn[1]), std::max(mx[2] - mn[2], 1.0e-6f));
  float sc = 1023.0f / ext;
  std::vector<std::pair<unsigned, int>> ki(n);
  for (int i = 0; i < n; i++) {
    unsigned ix = (unsigned)std::min(std::max((pts[i * 4] - mn[0]) * sc, 0.f), 1023.f);
    unsigned iy = (unsigned)std::min(std::max((pts[i * 4 + 1] - mn[1]) * sc, 0.f), 1023.f);
    unsigned iz = (unsigned)std::min(std::max((pts[i * 4 + 2] - mn[2]) * sc, 0.f), 1023.f);
    ki[i] = {hilbert30(ix, iy, iz), i};
  }
  std::sort(ki.begin(), ki.end());
  int nsup = (n + 1023) / 1024, nleaf = nsup * 32;
  H.sp.assign((size_t)nsup * 1024, make_float4(INFINITY, INFINITY, INFINITY, bits_idx(kPadIdx)));
  for (int s = 0; s < n; s++) { int i = ki[s].second; H.sp[s] = make_float4(pts[i * 4], pts[i * 4 + 1], pts[i * 4 + 2], bits_idx(i)); }
  H.llo.assign(nleaf, make_float4(INFINITY, INFINITY, INFINITY, 0)); H.lhi.assign(nleaf, make_float4(-INFINITY, -INFINITY, -INFINITY, 0));
  H.slo.assign(nsup, make_float4(INFINITY, INFINITY, INFINITY, 0)); H.shi.assign(nsup, make_float4(-INFINITY, -INFINITY, -INFINITY, 0));
  for (int s = 0; s < n; s++) {
    int l = s / 32, u = s / 1024;
    float4 p = H.sp[s];
    H.llo[l].x = std::min(H.llo[l].x, p.x); H.llo[l].y = std::min(H.llo[l].y, p.y); H.llo[l].z = std::min(H.llo[l].z, p.z);
    H.lhi[l].x = std::max(H.lhi[l].x, p.x); H.lhi[l].y = std::max(H.lhi[l].y, p.y); H.lhi[l].z = std::max(H.lhi[l].z, p.z);
    H.slo[u].x = std::min(H.slo[u].x, p.x); H.slo[u].y = std::min(H.slo[u].y, p.y); H.slo[u].z = std::min(H.slo[u].z, p.z);
    H.shi[u].x = std::max(H.shi[u].x, p.x); H.shi[u].y = std::max(H.shi[u].y, p.y); H.shi[u].z = std::max(H.shi[u].z, p.z);
  }
  H.b.sp = H.sp.data(); H.b.leaf_lo = H.llo.data(); H.b.leaf_hi = H.lhi.data(); H.b.sup_lo = H.slo.data(); H.b.sup_hi = H.shi.data();
  H.b.nleaf = nleaf; H.b.nsup = nsup; H.b.n = n;
  return H;
}

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

static double urand(unsigned long long& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return (double)(s >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 20000;
  int nq = argc > 2 ? atoi(argv[2]) : 5000;
  int mode = argc > 3 ? atoi(argv[3]) : 0;  // 0 lidar-ish, 1 lattice (many exact ties), 2 clustered + far outliers
  unsigned long long s = 12345 + mode;
  std::vector<float> pts(n * 4), qs(nq * 4);
  for (int i = 0; i < n; i++) {
    float x, y, z;
    if (mode == 1) { x = (float)((int)(urand(s) * 12)) * 0.5f; y = (float)((int)(urand(s) * 12)) * 0.5f; z = (float)((int)(urand(s) * 6)) * 0.25f; }
    else if (mode == 2) { double r = urand(s) < 0.98 ? 2.0 : 300.0; x = (float)((urand(s) - 0.5) * r); y = (float)((urand(s) - 0.5) * r); z = (float)((urand(s) - 0.5) * r * 0.2); }
    else { double a = urand(s) * 6.2831853, r = 1.0 + 60.0 * urand(s) * urand(s); x = (float)(r * cos(a)); y = (float)(r * sin(a)); z = (float)(-1.8 + 0.02 * urand(s) + (urand(s) < 0.2 ? 5 * urand(s) : 0)); }
    pts[i * 4] = x; pts[i * 4 + 1] = y; pts[i * 4 + 2] = z; pts[i * 4 + 3] = 1.f;
  }
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
