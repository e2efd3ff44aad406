// tests/leaf_obb.hpp — OFFLINE EXPERIMENT (round-2 candidate, not in the product): a per-leaf oriented bounding box (PCA frame of the
// leaf's 32 points) as a second-stage lower bound behind the AABB test of bvh_try_leaf (hook: -DB2R_LEAF_OBB).  LiDAR leaves are
// near-planar patches; an axis-aligned box of a tilted patch is fat, the oriented one is not.  Used by tools/warp_cost.cpp (visit
// counts) and tests/warp_harness.cpp (the bound stays conservative: results still equal the oracle's).
// Include once with LEAF_OBB_PART 1 before csrc/bvh.cuh and once with LEAF_OBB_PART 2 after tests/host_bvh.hpp.
#if LEAF_OBB_PART == 1
#include <vector>
#include <cmath>
// oriented box of a leaf: PCA frame (rows ax[0..2]) and the extent of the leaf's points along each axis, widened by a rounding slack
struct LeafObb { float ax[3][3]; float lo[3], hi[3]; bool valid; };
static const std::vector<LeafObb>* g_obb = nullptr;
static long g_obb_tests = 0, g_obb_rejects = 0;
static inline bool b2r_leaf_obb_pass(int l, float qx, float qy, float qz, float worst, float limit) {
  if (!g_obb) return true;
  const LeafObb& o = (*g_obb)[l];
  if (!o.valid) return true;
  g_obb_tests++;
  double lb = 0;
  for (int a = 0; a < 3; a++) {
    const float t = o.ax[a][0] * qx + o.ax[a][1] * qy + o.ax[a][2] * qz;
    float gap = 0.f;
    if (t < o.lo[a]) gap = o.lo[a] - t; else if (t > o.hi[a]) gap = t - o.hi[a];
    lb += (double)gap * gap;
  }
  const float lbf = (float)(lb * (1.0 - 1e-5));  // conservative: never above the true squared distance of any point of the leaf
  const bool pass = !(lbf > worst) && lbf < limit;
  if (!pass) g_obb_rejects++;
  return pass;
}
#elif LEAF_OBB_PART == 2
static std::vector<LeafObb> make_obbs(const HostBvh& H) {
    std::vector<LeafObb> v(H.b.nleaf);
    for (int l = 0; l < H.b.nleaf; l++) {
      LeafObb& o = v[l];
      o.valid = false;
      double m[3] = {0, 0, 0};
      int cnt = 0;
      for (int t = 0; t < kLeaf; t++) { const float4 p = H.sp[l * kLeaf + t]; if (idx_bits(p.w) == kPadIdx) continue; m[0] += p.x; m[1] += p.y; m[2] += p.z; cnt++; }
      if (cnt < 3) continue;
      for (int a = 0; a < 3; a++) m[a] /= cnt;
      double Cv[9] = {0};
      for (int t = 0; t < kLeaf; t++) {
        const float4 p = H.sp[l * kLeaf + t];
        if (idx_bits(p.w) == kPadIdx) continue;
        const double d[3] = {p.x - m[0], p.y - m[1], p.z - m[2]};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Cv[i * 3 + j] += d[i] * d[j];
      }
      // cyclic Jacobi: eigenvectors in the columns of V
      double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      for (int sweep = 0; sweep < 30; sweep++)
        for (int p = 0; p < 3; p++) for (int q = p + 1; q < 3; q++) {
          const double apq = Cv[p * 3 + q];
          if (std::fabs(apq) < 1e-300) continue;
          const double th = (Cv[q * 3 + q] - Cv[p * 3 + p]) / (2 * apq), tt = (th >= 0 ? 1 : -1) / (std::fabs(th) + std::sqrt(th * th + 1)), c = 1 / std::sqrt(tt * tt + 1), s2 = tt * c;
          for (int k = 0; k < 3; k++) { const double a = Cv[k * 3 + p], b2 = Cv[k * 3 + q]; Cv[k * 3 + p] = c * a - s2 * b2; Cv[k * 3 + q] = s2 * a + c * b2; }
          for (int k = 0; k < 3; k++) { const double a = Cv[p * 3 + k], b2 = Cv[q * 3 + k]; Cv[p * 3 + k] = c * a - s2 * b2; Cv[q * 3 + k] = s2 * a + c * b2; }
          for (int k = 0; k < 3; k++) { const double a = V[k * 3 + p], b2 = V[k * 3 + q]; V[k * 3 + p] = c * a - s2 * b2; V[k * 3 + q] = s2 * a + c * b2; }
        }
      for (int a = 0; a < 3; a++) { for (int k = 0; k < 3; k++) o.ax[a][k] = (float)V[k * 3 + a]; o.lo[a] = 1e30f; o.hi[a] = -1e30f; }
      for (int t = 0; t < kLeaf; t++) {
        const float4 p = H.sp[l * kLeaf + t];
        if (idx_bits(p.w) == kPadIdx) continue;
        for (int a = 0; a < 3; a++) { const float pr = o.ax[a][0] * p.x + o.ax[a][1] * p.y + o.ax[a][2] * p.z; o.lo[a] = std::min(o.lo[a], pr); o.hi[a] = std::max(o.hi[a], pr); }
      }
      for (int a = 0; a < 3; a++) { const float slack = 1e-4f + 4e-7f * (std::fabs(o.lo[a]) + std::fabs(o.hi[a])); o.lo[a] -= slack; o.hi[a] += slack; }
      o.valid = true;
    }
    return v;
  }
#endif
