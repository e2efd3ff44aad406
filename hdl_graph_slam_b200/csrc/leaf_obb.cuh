// leaf_obb.cuh — EXPERIMENTAL, compiled only with -DB2R_LEAF_OBB (the product build does not define it).
//
// A second-stage lower bound for bvh_try_leaf: the oriented box of a leaf's 32 points in their PCA frame.  LiDAR leaves are
// near-planar patches; the axis-aligned box of a tilted patch is fat, so a query that sits 1 m off a dense surface has to open every
// leaf of that surface within 1 m laterally, although only the one under its foot point can hold the answer
// (profiles/r01_h_offline_work_model.md: the oriented box rejects 42 % of the leaves a seeded 1-NN pass opens).
//
// Record per leaf (4 x float4):  r[i] = (axis_i.x, axis_i.y, axis_i.z, lo_i)  for i = 0..2,   r[3] = (hi_0, hi_1, hi_2, valid ? 1 : 0).
//
// Conservativeness (the bound must never exceed dist2_f32(q, p) of any point p of the leaf; ANY near-orthonormal frame gives a valid
// box, the PCA only makes it tight):
//   * extents are min / max of obb_proj(axis, p) over the leaf's points, evaluated with the SAME non-contracted float expression the
//     query uses, then widened by kObbRel * max_p |p|_1 ; the query subtracts kObbRel * |q|_1 from its gap.  obb_proj's rounding error is
//     below 3 * 2^-24 * |x|_1 (three products, two sums), kObbRel = 4 * 2^-24, so gap_i <= |axis_i . (q - p)| in real arithmetic;
//   * the axes are unit eigenvectors computed in float64 and rounded to float: | A A^T - I | < 1e-6, and dist2_f32 itself may round
//     down by a few 2^-24: the sum of the squared gaps is scaled by (1 - 4e-6).
// Pruning uses the same strict rule as the AABB bound (skip only if bound > worst), so ties stay reachable.
#pragma once
#include "common.cuh"
#include "linalg.cuh"

namespace b2r {

constexpr float kObbRel = 2.384185791015625e-7f;  // 4 * 2^-24

B2R_HD float obb_proj(float ax, float ay, float az, float x, float y, float z) { return fadd(fadd(fmul(ax, x), fmul(ay, y)), fmul(az, z)); }

// lower bound of dist2_f32(q, p) over the points p of the leaf described by r0..r3 (0 = no information)
B2R_HD float leaf_obb_bound2(const float4& r0, const float4& r1, const float4& r2, const float4& r3, float qx, float qy, float qz) {
  if (!(r3.w > 0.f)) return 0.f;
  const float sq = fmul(kObbRel, fadd(fadd(fabsf(qx), fabsf(qy)), fabsf(qz)));
  const float t0 = obb_proj(r0.x, r0.y, r0.z, qx, qy, qz), t1 = obb_proj(r1.x, r1.y, r1.z, qx, qy, qz), t2 = obb_proj(r2.x, r2.y, r2.z, qx, qy, qz);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (t0 < r0.w) g0 = fsub(fsub(r0.w, t0), sq); else if (t0 > r3.x) g0 = fsub(fsub(t0, r3.x), sq);
  if (t1 < r1.w) g1 = fsub(fsub(r1.w, t1), sq); else if (t1 > r3.y) g1 = fsub(fsub(t1, r3.y), sq);
  if (t2 < r2.w) g2 = fsub(fsub(r2.w, t2), sq); else if (t2 > r3.z) g2 = fsub(fsub(t2, r3.z), sq);
  g0 = fmaxf(g0, 0.f); g1 = fmaxf(g1, 0.f); g2 = fmaxf(g2, 0.f);
  return fmul(fadd(fadd(fmul(g0, g0), fmul(g1, g1)), fmul(g2, g2)), 0.999996f);
}

#ifdef B2R_WARP_CODE
// One warp = one leaf: lane t holds point t (valid = not padding).  Lane 0 writes the 4-float4 record to `out`.
__device__ __forceinline__ void leaf_obb_build_warp(float x, float y, float z, bool valid, float4* out) {
  const unsigned F = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int cnt = __popc(__ballot_sync(F, valid));
  if (cnt == 0) {
    if (lane == 0) { out[0] = out[1] = out[2] = make_float4(0.f, 0.f, 0.f, 0.f); out[3] = make_float4(0.f, 0.f, 0.f, 0.f); }
    return;
  }
  double s[3] = {valid ? (double)x : 0.0, valid ? (double)y : 0.0, valid ? (double)z : 0.0};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s[0] += __shfl_xor_sync(F, s[0], o); s[1] += __shfl_xor_sync(F, s[1], o); s[2] += __shfl_xor_sync(F, s[2], o); }
  const double inv = 1.0 / (double)cnt;
  const double mx = s[0] * inv, my = s[1] * inv, mz = s[2] * inv;
  const double dx = valid ? (double)x - mx : 0.0, dy = valid ? (double)y - my : 0.0, dz = valid ? (double)z - mz : 0.0;
  double c[6] = {dx * dx, dx * dy, dx * dz, dy * dy, dy * dz, dz * dz};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 6; i++) c[i] += __shfl_xor_sync(F, c[i], o);
  }
  // every lane holds the same sums (xor butterfly) => the same frame; no broadcast needed
  const double C[9] = {c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5]};
  double w[3], V[9];
  sym_eigen3(C, w, V);
  float a[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++) { a[i][0] = (float)V[0 * 3 + i]; a[i][1] = (float)V[1 * 3 + i]; a[i][2] = (float)V[2 * 3 + i]; }
  float lo[3], hi[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float t = obb_proj(a[i][0], a[i][1], a[i][2], x, y, z);
    lo[i] = valid ? t : INFINITY;
    hi[i] = valid ? t : -INFINITY;
  }
  float m1 = valid ? fadd(fadd(fabsf(x), fabsf(y)), fabsf(z)) : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 3; i++) { lo[i] = fminf(lo[i], __shfl_xor_sync(F, lo[i], o)); hi[i] = fmaxf(hi[i], __shfl_xor_sync(F, hi[i], o)); }
    m1 = fmaxf(m1, __shfl_xor_sync(F, m1, o));
  }
  const float wdn = fmul(kObbRel, m1);
  if (lane == 0) {
    out[0] = make_float4(a[0][0], a[0][1], a[0][2], fsub(lo[0], wdn));
    out[1] = make_float4(a[1][0], a[1][1], a[1][2], fsub(lo[1], wdn));
    out[2] = make_float4(a[2][0], a[2][1], a[2][2], fsub(lo[2], wdn));
    out[3] = make_float4(fadd(hi[0], wdn), fadd(hi[1], wdn), fadd(hi[2], wdn), 1.0f);
  }
}
#endif

}  // namespace b2r
