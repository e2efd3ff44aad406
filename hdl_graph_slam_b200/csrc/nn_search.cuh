// nn_search.cuh — exact nearest-neighbour traversal of the uniform grid (shared by the 1-NN correspondence search,
// the k-NN covariance kernel and the fitness kernel).
//
// Replaces the FLANN kd-tree queries the reference's registration handle performs (fast_gicp update_correspondences /
// calculate_covariances, pcl::Registration::getFitnessScore; call sites apps/scan_matching_odometry_nodelet.cpp:210,307,316,
// include/hdl_graph_slam/loop_detector.hpp:143,146).  Result contract (oracle/kdtree.hpp): the lexicographically smallest
// (d2, original index) with d2 evaluated in non-contracted float32.
//
// Exactness argument: cells are exact (Grid comment), so every point of an unvisited cell lies beyond the cell's face;
// float subtraction, multiplication and addition of non-negatives are monotone under round-to-nearest, hence a bound
// assembled from face distances in the same operation order never exceeds the d2 the point would evaluate to.
// Rows/shells are skipped only when bound > worst (strict), so equal-distance lower-index candidates stay reachable.
#pragma once
#include "common.cuh"

namespace b2r {

// axis lower bound between q and the half-open interval [lo, hi)
B2R_HD float axis_bound(float q, float lo, float hi) {
  if (q < lo) return fsub(lo, q);
  if (q > hi) return fsub(q, hi);
  return 0.f;
}

template <class Visitor>
B2R_HD void grid_scan_run(const int* __restrict__ cell_start, const float4* __restrict__ sp, int cell_a, int cell_b_incl,
                          float qx, float qy, float qz, Visitor& v) {
  int s = cell_start[cell_a];
  int e = cell_start[cell_b_incl + 1];
  for (int i = s; i < e; i++) {
    float4 p = sp[i];
    float d2 = dist2_f32(qx, qy, qz, p.x, p.y, p.z);
    v.visit(d2, idx_bits(p.w), i);
  }
}

// Visitor interface:  float worst() const;  float limit() const;  void visit(float d2, int orig_idx, int pos);
template <class Visitor>
B2R_HD void grid_search(const Grid& g, const int* __restrict__ cell_start, const float4* __restrict__ sp, float qx, float qy,
                        float qz, Visitor& v) {
  if (g.n_valid <= 0) return;
  const int cx = cell_coord(qx, g.ox, g.inv_h, g.nx);
  const int cy = cell_coord(qy, g.oy, g.inv_h, g.ny);
  const int cz = cell_coord(qz, g.oz, g.inv_h, g.nz);
  const float h = g.h;
  for (int r = 0;; r++) {
    const int x0 = cx - r, x1 = cx + r, y0 = cy - r, y1 = cy + r, z0 = cz - r, z1 = cz + r;
    const int xa = x0 < 0 ? 0 : x0, xb = x1 > g.nx - 1 ? g.nx - 1 : x1;
    const int ya = y0 < 0 ? 0 : y0, yb = y1 > g.ny - 1 ? g.ny - 1 : y1;
    const int za = z0 < 0 ? 0 : z0, zb = z1 > g.nz - 1 ? g.nz - 1 : z1;
    // x bound of the full-width run and of the two end cells
    const float bx_run = axis_bound(qx, fadd(g.ox, fmul((float)xa, h)), fadd(g.ox, fmul((float)(xb + 1), h)));
    const float bx_lo = (x0 >= 0) ? axis_bound(qx, fadd(g.ox, fmul((float)x0, h)), fadd(g.ox, fmul((float)(x0 + 1), h))) : 0.f;
    const float bx_hi = (x1 <= g.nx - 1) ? axis_bound(qx, fadd(g.ox, fmul((float)x1, h)), fadd(g.ox, fmul((float)(x1 + 1), h))) : 0.f;
    for (int z = za; z <= zb; z++) {
      const float bz = axis_bound(qz, fadd(g.oz, fmul((float)z, h)), fadd(g.oz, fmul((float)(z + 1), h)));
      const float bz2 = fmul(bz, bz);
      const bool zedge = (z == z0) || (z == z1);
      for (int y = ya; y <= yb; y++) {
        const float by = axis_bound(qy, fadd(g.oy, fmul((float)y, h)), fadd(g.oy, fmul((float)(y + 1), h)));
        const float by2 = fmul(by, by);
        const int row = (z * g.ny + y) * g.nx;
        if (zedge || y == y0 || y == y1) {
          float b = fadd(fadd(fmul(bx_run, bx_run), by2), bz2);
          if (!(b > v.worst())) grid_scan_run(cell_start, sp, row + xa, row + xb, qx, qy, qz, v);
        } else {
          if (x0 >= 0) {
            float b = fadd(fadd(fmul(bx_lo, bx_lo), by2), bz2);
            if (!(b > v.worst())) grid_scan_run(cell_start, sp, row + x0, row + x0, qx, qy, qz, v);
          }
          if (x1 <= g.nx - 1 && r > 0) {
            float b = fadd(fadd(fmul(bx_hi, bx_hi), by2), bz2);
            if (!(b > v.worst())) grid_scan_run(cell_start, sp, row + x1, row + x1, qx, qy, qz, v);
          }
        }
      }
    }
    // termination: distance to the nearest block face that still has grid behind it
    float fb = INFINITY;
    if (x0 > 0) fb = fminf(fb, fsub(qx, fadd(g.ox, fmul((float)x0, h))));
    if (x1 < g.nx - 1) fb = fminf(fb, fsub(fadd(g.ox, fmul((float)(x1 + 1), h)), qx));
    if (y0 > 0) fb = fminf(fb, fsub(qy, fadd(g.oy, fmul((float)y0, h))));
    if (y1 < g.ny - 1) fb = fminf(fb, fsub(fadd(g.oy, fmul((float)(y1 + 1), h)), qy));
    if (z0 > 0) fb = fminf(fb, fsub(qz, fadd(g.oz, fmul((float)z0, h))));
    if (z1 < g.nz - 1) fb = fminf(fb, fsub(fadd(g.oz, fmul((float)(z1 + 1), h)), qz));
    if (fb == INFINITY) break;  // block covers the whole grid
    if (fb < 0.f) fb = 0.f;     // cannot happen for finite q (q is clamped into the block); defensive
    const float fb2 = fmul(fb, fb);
    if (v.worst() < fb2) break;      // every unvisited point evaluates to d2 >= fb2 > worst
    if (!(fb2 < v.limit())) break;   // every unvisited point is at or beyond the caller's range limit
  }
}

// 1-NN visitor
struct Nn1 {
  float best_d2;
  int best_idx;
  int best_pos;
  float lim;
  B2R_HD float worst() const { return best_d2; }
  B2R_HD float limit() const { return lim; }
  B2R_HD void visit(float d2, int idx, int pos) {
    if (d2 < best_d2 || (d2 == best_d2 && idx < best_idx)) {
      best_d2 = d2;
      best_idx = idx;
      best_pos = pos;
    }
  }
};

}  // namespace b2r
