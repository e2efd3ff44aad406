// tests/host_bvh.hpp — serial host construction of the engine's implicit BVH (same rules as k_morton_keys / k_bvh_leaves) and the
// synthetic point sets shared by the CPU-side harnesses.  Test infrastructure only.
#pragma once
#include <vector>
#include <algorithm>
#include <cmath>

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


static double urand(unsigned long long& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return (double)(s >> 11) / 9007199254740992.0;
}

// mode 0 lidar-ish, 1 lattice (many exact ties), 2 clustered + far outliers
static std::vector<float> make_points(int n, int mode, unsigned long long& s) {
  std::vector<float> pts((size_t)n * 4);
  for (int i = 0; i < n; i++) {
    float x, y, z;
    if (mode == 1) { x = (float)((int)(urand(s) * 12)) * 0.5f; y = (float)((int)(urand(s) * 12)) * 0.5f; z = (float)((int)(urand(s) * 6)) * 0.25f; }
    else if (mode == 2) { double r = urand(s) < 0.98 ? 2.0 : 300.0; x = (float)((urand(s) - 0.5) * r); y = (float)((urand(s) - 0.5) * r); z = (float)((urand(s) - 0.5) * r * 0.2); }
    else { double a = urand(s) * 6.2831853, r = 1.0 + 60.0 * urand(s) * urand(s); x = (float)(r * cos(a)); y = (float)(r * sin(a)); z = (float)(-1.8 + 0.02 * urand(s) + (urand(s) < 0.2 ? 5 * urand(s) : 0)); }
    pts[i * 4] = x; pts[i * 4 + 1] = y; pts[i * 4 + 2] = z; pts[i * 4 + 3] = 1.f;
  }
  return pts;
}
