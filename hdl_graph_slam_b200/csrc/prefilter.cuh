// prefilter.cuh — the rest of the prefilter chain next to the voxel grid ("next" rows):
//   k_distance_flags   <- PrefilteringNodelet::distance_filter          /root/reference/apps/prefiltering_nodelet.cpp:164-180
//   k_knn_stat         <- pcl::RadiusOutlierRemoval / pcl::StatisticalOutlierRemoval as configured at :72-93 and run at :151-162
// Both outlier filters are nearest-neighbour statistics of the cloud against itself, so they reuse the BVH k-NN traversal:
//   RADIUS      keep p  iff  its (min_neighbors+1)-th nearest neighbour (itself included) lies within the radius
//   STATISTICAL keep p  iff  mean distance to its mean_k nearest neighbours (itself excluded) <= mean + mul * stddev over the cloud
// The device produces one float per point (k-th squared distance, or mean neighbour distance); the O(n) decision and the
// order-preserving compaction are done on the host exactly as PCL does them (double accumulators, index order).
#pragma once
#include "common.cuh"
#include "bvh.cuh"
#include "gicp.cuh"

namespace b2r {

__global__ void k_distance_flags(const float* __restrict__ raw, int stride_f, int n, double near_t, double far_t, unsigned char* flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = raw + (size_t)i * stride_f;
  const float x = p[0], y = p[1], z = p[2];
  // Eigen Vector3f::norm(): sqrt((x*x + y*y) + z*z) in float32, then promoted to double for the comparison
  const float sq = fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z));
  const double d = (double)__fsqrt_rn(sq);
  flags[i] = (finite3(x, y, z) && d > near_t && d < far_t) ? 1 : 0;
}

// mode 0: out[orig idx] = squared distance of the k-th nearest neighbour (self included), +inf if fewer than k points
// mode 1: out[orig idx] = (float)( sum_{j=1..k-1} sqrtf(d2_j) / (k-1) )   (PCL StatisticalOutlierRemoval, k = mean_k + 1)
__global__ void __launch_bounds__(kKnnThreads, 4) k_knn_stat(Bvh b, int k, int mode, float* __restrict__ out) {
  extern __shared__ unsigned long long knn_keys[];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int leaf = s >> 5;
  if (leaf >= b.nleaf) return;
  const float4 q = b.sp[s];
  const bool active = idx_bits(q.w) != kPadIdx;
  KnnList L;
  L.key = knn_keys + threadIdx.x;
  L.k = k;
  L.cnt = 0;
  L.stride = blockDim.x;
  L.wkey = kKeyInf;
  bvh_group_search(b, q.x, q.y, q.z, active, L, leaf);
  if (!active) return;
  float v;
  if (mode == 0) {
    v = (L.cnt == k) ? nn_key_d2(L.key[(k - 1) * L.stride]) : INFINITY;
  } else {
    double sum = 0.0;
    for (int j = 1; j < L.cnt; j++) sum += (double)__fsqrt_rn(nn_key_d2(L.key[j * L.stride]));
    v = (float)(sum / (double)(k - 1));
  }
  out[idx_bits(q.w)] = v;
}

}  // namespace b2r
