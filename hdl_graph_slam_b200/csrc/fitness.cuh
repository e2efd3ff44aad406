// fitness.cuh — nearest-neighbour fitness score, inlier count, exact 1-NN queries and the `aligned` output cloud.
//
// k_fitness      <- pcl::Registration::getFitnessScore(max_range) (called at apps/scan_matching_odometry_nodelet.cpp:307,
//                   include/hdl_graph_slam/loop_detector.hpp:146; in-tree twin src/hdl_graph_slam/information_matrix_calculator.cpp:49-80)
//                   fused with the inlier loop of apps/scan_matching_odometry_nodelet.cpp:309-320
// k_nearest      <- getSearchMethodTarget()->nearestKSearch(pt, 1, ...) (apps/scan_matching_odometry_nodelet.cpp:316)
// k_transform    <- the `aligned` cloud written by align() (pcl::transformPointCloud with final_transformation_)
#pragma once
#include "common.cuh"
#include "bvh.cuh"
#include "gicp.cuh"

namespace b2r {

struct FitArgs {
  Bvh src;                // queries = the source's sorted points (a warp = one leaf = spatial neighbours)
  Bvh tgt;
  float Tf[12];
  double max_range;       // compared with the SQUARED distance (reference semantics)
  float inlier_thresh_sq;
  double* partials;       // [blocks][3]
  double* out;            // [3] = sum d2, count, inliers
  unsigned int* counter;
};

__global__ void __launch_bounds__(kLinThreads, 4) k_fitness(const __grid_constant__ FitArgs A) {
  __shared__ double red[3 * 32];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[3] = {0.0, 0.0, 0.0};
  float4 p = make_float4(0.f, 0.f, 0.f, bits_idx(kPadIdx));
  if (s < A.src.nleaf * kLeaf) p = A.src.sp[s];
  float qx = 0.f, qy = 0.f, qz = 0.f;
  Nn1 v;
  v.reset(INFINITY);
  bool active = false;
  if (idx_bits(p.w) != kPadIdx) {
    qx = xform_row(A.Tf[0], A.Tf[1], A.Tf[2], A.Tf[3], p.x, p.y, p.z);
    qy = xform_row(A.Tf[4], A.Tf[5], A.Tf[6], A.Tf[7], p.x, p.y, p.z);
    qz = xform_row(A.Tf[8], A.Tf[9], A.Tf[10], A.Tf[11], p.x, p.y, p.z);
    active = finite3(qx, qy, qz);
  }
  bvh_group_search(A.tgt, qx, qy, qz, active, v, -1);
  if (active && v.best_pos >= 0) {
    if ((double)v.best_d2() <= A.max_range) { acc[0] = (double)v.best_d2(); acc[1] = 1.0; }
    if (v.best_d2() < A.inlier_thresh_sq) acc[2] = 1.0;
  }
  block_reduce<3>(acc, red);
  finish_partials<3>(acc, A.partials, A.out, A.counter);
}

// queries in caller order (no spatial coherence guaranteed: still exact, only less efficient)
__global__ void __launch_bounds__(256, 2) k_nearest(const float* __restrict__ q_raw, int stride_f, int n, const __grid_constant__ Bvh tgt,
                                                    int* idx_out, float* d2_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  Nn1 v;
  v.reset(INFINITY);
  bool active = false;
  if (i < n) {
    const float* p = q_raw + (size_t)i * stride_f;
    qx = p[0]; qy = p[1]; qz = p[2];
    active = finite3(qx, qy, qz);
  }
  bvh_group_search(tgt, qx, qy, qz, active, v, -1);
  if (i < n) {
    idx_out[i] = v.best_pos >= 0 ? v.best_idx() : -1;
    d2_out[i] = v.best_d2();
  }
}

struct XfArg { float Tf[12]; };
__global__ void k_transform(const float* __restrict__ raw, int stride_f, int n, XfArg X, float4* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = raw + (size_t)i * stride_f;
  const float x = p[0], y = p[1], z = p[2];
  out[i] = make_float4(xform_row(X.Tf[0], X.Tf[1], X.Tf[2], X.Tf[3], x, y, z), xform_row(X.Tf[4], X.Tf[5], X.Tf[6], X.Tf[7], x, y, z),
                       xform_row(X.Tf[8], X.Tf[9], X.Tf[10], X.Tf[11], x, y, z), 1.0f);
}

}  // namespace b2r
