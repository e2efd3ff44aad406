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
//
// Two regimes (profiles/r01_a: a purely thread-local shell walk has a catastrophic latency tail on far-field LiDAR points):
//   grid_search(..., max_r)   thread-local shells r = 0..max_r  — finishes the dense majority of queries
//   warp_box_scan / warp_finish_nn1   the 32 lanes of a warp split the rows of a bounded box for ONE unfinished query
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

// squared distance from q to the nearest face of the block [c-R, c+R]^3 that still has grid behind it;
// returns INFINITY when the block covers the whole grid.
B2R_HD float block_face_bound2(const Grid& g, float qx, float qy, float qz, int cx, int cy, int cz, int R) {
  const float h = g.h;
  const long long x0 = (long long)cx - R, x1 = (long long)cx + R, y0 = (long long)cy - R, y1 = (long long)cy + R, z0 = (long long)cz - R, z1 = (long long)cz + R;
  float fb = INFINITY;
  if (x0 > 0) fb = fminf(fb, fsub(qx, fadd(g.ox, fmul((float)x0, h))));
  if (x1 < g.nx - 1) fb = fminf(fb, fsub(fadd(g.ox, fmul((float)(x1 + 1), h)), qx));
  if (y0 > 0) fb = fminf(fb, fsub(qy, fadd(g.oy, fmul((float)y0, h))));
  if (y1 < g.ny - 1) fb = fminf(fb, fsub(fadd(g.oy, fmul((float)(y1 + 1), h)), qy));
  if (z0 > 0) fb = fminf(fb, fsub(qz, fadd(g.oz, fmul((float)z0, h))));
  if (z1 < g.nz - 1) fb = fminf(fb, fsub(fadd(g.oz, fmul((float)(z1 + 1), h)), qz));
  if (fb == INFINITY) return INFINITY;
  if (fb < 0.f) fb = 0.f;  // cannot happen for finite q (q is clamped into the block); defensive
  return fmul(fb, fb);
}

// Visitor interface:  float worst() const;  float limit() const;  void visit(float d2, int orig_idx, int pos);
// Returns true when the search is complete (result exact), false when it stopped at max_r with work left.
template <class Visitor>
B2R_HD bool grid_search(const Grid& g, const int* __restrict__ cell_start, const float4* __restrict__ sp, float qx, float qy,
                        float qz, Visitor& v, int max_r = 0x3fffffff) {
  if (g.n_valid <= 0) return true;
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
    const float fb2 = block_face_bound2(g, qx, qy, qz, cx, cy, cz, r);
    if (fb2 == INFINITY) return true;      // block covers the whole grid
    if (v.worst() < fb2) return true;      // every unvisited point evaluates to d2 >= fb2 > worst
    if (!(fb2 < v.limit())) return true;   // every unvisited point is at or beyond the caller's range limit
    if (r >= max_r) return false;
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

#ifdef __CUDACC__
// The calling warp scans the box [c-R, c+R]^3 (clipped) for ONE query: lane l takes (z,y) rows l, l+32, ...; each row is
// one contiguous run of the sorted array.  Every lane accumulates into its own visitor.
template <class Visitor>
__device__ __forceinline__ void warp_box_scan(const Grid& g, const int* __restrict__ cell_start, const float4* __restrict__ sp, float qx,
                                              float qy, float qz, int cx, int cy, int cz, int R, int lane, Visitor& v) {
  const float h = g.h;
  const int xa = (cx - R < 0 || R > g.nx) ? 0 : cx - R, xb = (R > g.nx || cx + R > g.nx - 1) ? g.nx - 1 : cx + R;
  const int ya = (cy - R < 0 || R > g.ny) ? 0 : cy - R, yb = (R > g.ny || cy + R > g.ny - 1) ? g.ny - 1 : cy + R;
  const int za = (cz - R < 0 || R > g.nz) ? 0 : cz - R, zb = (R > g.nz || cz + R > g.nz - 1) ? g.nz - 1 : cz + R;
  const int ny_r = yb - ya + 1;
  const int total = ny_r * (zb - za + 1);
  const float bx = axis_bound(qx, fadd(g.ox, fmul((float)xa, h)), fadd(g.ox, fmul((float)(xb + 1), h)));
  const float bx2 = fmul(bx, bx);
  for (int t = lane; t < total; t += 32) {
    const int z = za + t / ny_r, y = ya + t % ny_r;
    const float by = axis_bound(qy, fadd(g.oy, fmul((float)y, h)), fadd(g.oy, fmul((float)(y + 1), h)));
    const float bz = axis_bound(qz, fadd(g.oz, fmul((float)z, h)), fadd(g.oz, fmul((float)(z + 1), h)));
    const float b = fadd(fadd(bx2, fmul(by, by)), fmul(bz, bz));
    if (b > fminf(v.worst(), v.limit())) continue;  // rows at or beyond the caller's range limit cannot matter
    const int row = (z * g.ny + y) * g.nx;
    grid_scan_run(cell_start, sp, row + xa, row + xb, qx, qy, qz, v);
  }
}

// ---- warp-group search ("shared staging of cells"): the 32 lanes of a warp hold 32 queries that are neighbours in the
// sorted order, i.e. sit in the same or adjacent cells.  The warp walks the rings around the GROUP's cell box together:
// every candidate run is loaded once, coalesced (lane j loads point base+j), and broadcast by shuffles so that each lane
// tests it against its own query.  This removes the SIMT divergence of 32 independent shell walks (profiles/r01_b).
// Per-lane exactness is unchanged: a lane skips a run only if the run's box bound exceeds its own worst (strict), and is
// done only when its worst is below the distance to the group box's faces.
// Returns false (nothing done) when the group's cells are too spread out to share work; done[lane] tells completion.
constexpr int kGroupMaxDX = 5, kGroupMaxDY = 1, kGroupMaxDZ = 1;

template <class Visitor>
__device__ __forceinline__ bool warp_group_search(const Grid& g, const int* __restrict__ cell_start, const float4* __restrict__ sp, float qx,
                                                  float qy, float qz, bool active, Visitor& v, int max_r, bool& done) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  done = !active;
  if (g.n_valid <= 0) { done = true; return true; }
  const int cx = active ? cell_coord(qx, g.ox, g.inv_h, g.nx) : 0;
  const int cy = active ? cell_coord(qy, g.oy, g.inv_h, g.ny) : 0;
  const int cz = active ? cell_coord(qz, g.oz, g.inv_h, g.nz) : 0;
  int gx0 = active ? cx : 0x7fffffff, gx1 = active ? cx : -1, gy0 = active ? cy : 0x7fffffff, gy1 = active ? cy : -1;
  int gz0 = active ? cz : 0x7fffffff, gz1 = active ? cz : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gx0 = min(gx0, __shfl_xor_sync(FULL, gx0, o)); gx1 = max(gx1, __shfl_xor_sync(FULL, gx1, o));
    gy0 = min(gy0, __shfl_xor_sync(FULL, gy0, o)); gy1 = max(gy1, __shfl_xor_sync(FULL, gy1, o));
    gz0 = min(gz0, __shfl_xor_sync(FULL, gz0, o)); gz1 = max(gz1, __shfl_xor_sync(FULL, gz1, o));
  }
  if (gx1 < 0) return true;  // no active lane
  if (gx1 - gx0 > kGroupMaxDX || gy1 - gy0 > kGroupMaxDY || gz1 - gz0 > kGroupMaxDZ) return false;
  const float h = g.h;
  for (int r = 0; r <= max_r; r++) {
    const int X0 = gx0 - r, X1 = gx1 + r, Y0 = gy0 - r, Y1 = gy1 + r, Z0 = gz0 - r, Z1 = gz1 + r;
    const int xa = X0 < 0 ? 0 : X0, xb = X1 > g.nx - 1 ? g.nx - 1 : X1;
    const int ya = Y0 < 0 ? 0 : Y0, yb = Y1 > g.ny - 1 ? g.ny - 1 : Y1;
    const int za = Z0 < 0 ? 0 : Z0, zb = Z1 > g.nz - 1 ? g.nz - 1 : Z1;
    for (int z = za; z <= zb; z++) {
      const float bz = axis_bound(qz, fadd(g.oz, fmul((float)z, h)), fadd(g.oz, fmul((float)(z + 1), h)));
      const float bz2 = fmul(bz, bz);
      for (int y = ya; y <= yb; y++) {
        const float by = axis_bound(qy, fadd(g.oy, fmul((float)y, h)), fadd(g.oy, fmul((float)(y + 1), h)));
        const float by2 = fmul(by, by);
        const bool outer = (r == 0) || z == Z0 || z == Z1 || y == Y0 || y == Y1;
        const int row = (z * g.ny + y) * g.nx;
        // an outer row contributes its full x-run; an inner row only the two cells the ring adds at its ends
        for (int part = 0; part < (outer ? 1 : 2); part++) {
          int ra, rb;
          if (outer) { ra = xa; rb = xb; }
          else if (part == 0) { if (X0 < 0) continue; ra = rb = X0; }
          else { if (X1 > g.nx - 1) continue; ra = rb = X1; }
          const float bx = axis_bound(qx, fadd(g.ox, fmul((float)ra, h)), fadd(g.ox, fmul((float)(rb + 1), h)));
          const float b = fadd(fadd(fmul(bx, bx), by2), bz2);  // same association as dist2_f32 => monotone lower bound
          const bool skip = done || (b > v.worst());
          if (__ballot_sync(FULL, !skip) == 0) continue;
          const int s = cell_start[row + ra], e = cell_start[row + rb + 1];
          for (int base = s; base < e; base += 32) {
            const int mine = base + lane;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mine < e) p = sp[mine];
            const int cnt = min(32, e - base);
            for (int t = 0; t < cnt; t++) {
              const float px = __shfl_sync(FULL, p.x, t), py = __shfl_sync(FULL, p.y, t), pz = __shfl_sync(FULL, p.z, t);
              const float pw = __shfl_sync(FULL, p.w, t);
              if (!skip) v.visit(dist2_f32(qx, qy, qz, px, py, pz), idx_bits(pw), base + t);
            }
          }
        }
      }
    }
    // per-lane termination against the faces of the group box (only faces with grid behind them)
    if (!done) {
      float fb = INFINITY;
      if (X0 > 0) fb = fminf(fb, fsub(qx, fadd(g.ox, fmul((float)X0, h))));
      if (X1 < g.nx - 1) fb = fminf(fb, fsub(fadd(g.ox, fmul((float)(X1 + 1), h)), qx));
      if (Y0 > 0) fb = fminf(fb, fsub(qy, fadd(g.oy, fmul((float)Y0, h))));
      if (Y1 < g.ny - 1) fb = fminf(fb, fsub(fadd(g.oy, fmul((float)(Y1 + 1), h)), qy));
      if (Z0 > 0) fb = fminf(fb, fsub(qz, fadd(g.oz, fmul((float)Z0, h))));
      if (Z1 < g.nz - 1) fb = fminf(fb, fsub(fadd(g.oz, fmul((float)(Z1 + 1), h)), qz));
      if (fb == INFINITY) done = true;
      else {
        if (fb < 0.f) fb = 0.f;
        const float fb2 = fmul(fb, fb);
        if (v.worst() < fb2 || !(fb2 < v.limit())) done = true;
      }
    }
    if (__ballot_sync(FULL, !done) == 0) break;
  }
  return true;
}

__device__ __forceinline__ void warp_min_nn1(Nn1& v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, v.best_d2, o);
    const int oi = __shfl_xor_sync(0xffffffffu, v.best_idx, o);
    const int op = __shfl_xor_sync(0xffffffffu, v.best_pos, o);
    if (od < v.best_d2 || (od == v.best_d2 && oi < v.best_idx)) { v.best_d2 = od; v.best_idx = oi; v.best_pos = op; }
  }
}

// Must be called by ALL 32 lanes of a warp.  Lanes with need == true hold an unfinished 1-NN query (qx,qy,qz, v after the
// thread-local phase); the warp finishes them one at a time with growing boxes.  Exact (same argument as grid_search).
__device__ __forceinline__ void warp_finish_nn1(const Grid& g, const int* __restrict__ cell_start, const float4* __restrict__ sp, float qx,
                                                float qy, float qz, Nn1& v, bool need) {
  const int lane = threadIdx.x & 31;
  unsigned hard = __ballot_sync(0xffffffffu, need);
  while (hard) {
    const int L = __ffs(hard) - 1;
    hard &= hard - 1;
    const float ax = __shfl_sync(0xffffffffu, qx, L), ay = __shfl_sync(0xffffffffu, qy, L), az = __shfl_sync(0xffffffffu, qz, L);
    Nn1 w;
    w.best_d2 = __shfl_sync(0xffffffffu, v.best_d2, L);
    w.best_idx = __shfl_sync(0xffffffffu, v.best_idx, L);
    w.best_pos = __shfl_sync(0xffffffffu, v.best_pos, L);
    w.lim = __shfl_sync(0xffffffffu, v.lim, L);
    const int cx = cell_coord(ax, g.ox, g.inv_h, g.nx), cy = cell_coord(ay, g.oy, g.inv_h, g.ny), cz = cell_coord(az, g.oz, g.inv_h, g.nz);
    // first box: large enough for the current bound (best found so far, or the caller's range limit), at least 4 cells
    float bound = fminf(w.best_d2, w.lim);
    int R = 4;
    if (bound < 1.0e30f) {
      const float rr = sqrtf(bound) * g.inv_h;
      R = rr < 1.0e6f ? (int)rr + 2 : 0x3fffffff;
    }
    for (;;) {
      warp_box_scan(g, cell_start, sp, ax, ay, az, cx, cy, cz, R, lane, w);
      warp_min_nn1(w);
      const float fb2 = block_face_bound2(g, ax, ay, az, cx, cy, cz, R);
      if (fb2 == INFINITY || w.best_d2 < fb2 || !(fb2 < w.lim)) break;
      R = (R < (1 << 20)) ? R * 4 : 0x3fffffff;
    }
    if (lane == L) { v.best_d2 = w.best_d2; v.best_idx = w.best_idx; v.best_pos = w.best_pos; }
  }
}
#endif

}  // namespace b2r
