// tests/host_harness.cu — CPU-side check of the engine's exact grid search LOGIC (nn_search.cuh compiled as host code)
// against the oracle's kd-tree.  Test infrastructure: built by tests/test_host_logic.py with nvcc, never shipped.
// It mirrors grid.cuh's build serially on the host (same formulas) and then runs grid_search for 1-NN and k-NN.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>
#include "../hdl_graph_slam_b200/csrc/common.cuh"
#include "../hdl_graph_slam_b200/csrc/nn_search.cuh"
#include "../oracle/oracle.h"

using namespace b2r;

struct HostGrid {
  Grid g;
  std::vector<int> cell_start;
  std::vector<float4> sorted;
};

static HostGrid build(const std::vector<float>& pts, int n, float h_min, int cap) {
  HostGrid G;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], pts[i * 4 + d]); mx[d] = std::max(mx[d], pts[i * 4 + d]); }
  float h = h_min;
  for (;;) {
    float inv = 1.f / h;
    double cells = 1;
    int dims[3];
    float o[3];
    for (int d = 0; d < 3; d++) {
      o[d] = floorf(mn[d] * inv) * h;
      dims[d] = (int)floorf((mx[d] - o[d]) * inv) + 1;
      cells *= dims[d];
    }
    if (cells <= cap) {
      G.g.ox = o[0]; G.g.oy = o[1]; G.g.oz = o[2]; G.g.h = h; G.g.inv_h = inv;
      G.g.nx = dims[0]; G.g.ny = dims[1]; G.g.nz = dims[2]; G.g.ncell = dims[0] * dims[1] * dims[2];
      break;
    }
    h *= 2.f;
  }
  G.g.n = n; G.g.n_valid = n;
  std::vector<std::pair<int, int>> ci(n);
  for (int i = 0; i < n; i++) {
    int cx = cell_coord(pts[i * 4 + 0], G.g.ox, G.g.inv_h, G.g.nx);
    int cy = cell_coord(pts[i * 4 + 1], G.g.oy, G.g.inv_h, G.g.ny);
    int cz = cell_coord(pts[i * 4 + 2], G.g.oz, G.g.inv_h, G.g.nz);
    ci[i] = {(cz * G.g.ny + cy) * G.g.nx + cx, i};
  }
  std::sort(ci.begin(), ci.end());
  G.cell_start.assign(G.g.ncell + 1, 0);
  for (auto& c : ci) G.cell_start[c.first + 1]++;
  for (int c = 0; c < G.g.ncell; c++) G.cell_start[c + 1] += G.cell_start[c];
  G.sorted.resize(n);
  for (int s = 0; s < n; s++) {
    int i = ci[s].second;
    G.sorted[s] = make_float4(pts[i * 4], pts[i * 4 + 1], pts[i * 4 + 2], bits_idx(i));
  }
  return G;
}

struct HostKnn {
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
  float h_min = argc > 4 ? (float)atof(argv[4]) : 0.5f;
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
    if (i % 97 == 0) { qs[i * 4] += 400.f; }  // far outside the grid
  }
  HostGrid G = build(pts, n, h_min, 1 << 22);
  printf("grid h=%g dims=%d %d %d ncell=%d\n", G.g.h, G.g.nx, G.g.ny, G.g.nz, G.g.ncell);
  const int k = 20;
  std::vector<int32_t> oidx((size_t)nq * k);
  std::vector<float> od2((size_t)nq * k);
  orc_knn(pts.data(), n, 4, qs.data(), nq, 4, k, oidx.data(), od2.data(), 0);
  long bad1 = 0, badk = 0, badlim = 0;
  for (int i = 0; i < nq; i++) {
    Nn1 v; v.best_d2 = INFINITY; v.best_idx = 0x7fffffff; v.best_pos = -1; v.lim = INFINITY;
    grid_search(G.g, G.cell_start.data(), G.sorted.data(), qs[i * 4], qs[i * 4 + 1], qs[i * 4 + 2], v);
    if (v.best_idx != oidx[(size_t)i * k] || v.best_d2 != od2[(size_t)i * k]) {
      if (bad1 < 5) printf("1nn mismatch q%d: got (%g,%d) want (%g,%d)\n", i, v.best_d2, v.best_idx, od2[(size_t)i * k], oidx[(size_t)i * k]);
      bad1++;
    }
    // range-limited search (GICP: limit 6.25): result must agree whenever the true NN is inside the limit
    Nn1 w; w.best_d2 = INFINITY; w.best_idx = 0x7fffffff; w.best_pos = -1; w.lim = 6.25f;
    grid_search(G.g, G.cell_start.data(), G.sorted.data(), qs[i * 4], qs[i * 4 + 1], qs[i * 4 + 2], w);
    bool want_valid = od2[(size_t)i * k] < 6.25f;
    bool got_valid = w.best_pos >= 0 && w.best_d2 < 6.25f;
    if (want_valid != got_valid || (want_valid && (w.best_idx != oidx[(size_t)i * k] || w.best_d2 != od2[(size_t)i * k]))) badlim++;
    HostKnn K; K.k = k; K.d.resize(k); K.id.resize(k);
    grid_search(G.g, G.cell_start.data(), G.sorted.data(), qs[i * 4], qs[i * 4 + 1], qs[i * 4 + 2], K);
    for (int j = 0; j < k; j++)
      if (K.id[j] != oidx[(size_t)i * k + j] || K.d[j] != od2[(size_t)i * k + j]) { badk++; break; }
  }
  printf("queries=%d  1nn_mismatch=%ld  limited_mismatch=%ld  knn_mismatch=%ld\n", nq, bad1, badlim, badk);
  return (bad1 || badk || badlim) ? 1 : 0;
}
