// map_cloud.cuh — MapCloudGenerator::generate on the device ("next" row).
//
// Reference: /root/reference/src/hdl_graph_slam/map_cloud_generator.cpp:13-51 (called from apps/hdl_graph_slam_nodelet.cpp:528,989):
//   for every keyframe snapshot: dst = pose.cast<float>() * src (Vector4f, w = 1), intensity copied, all keyframes concatenated;
//   resolution <= 0: that cloud is the result; otherwise pcl::octree::OctreePointCloud(resolution).addPointsFromInputCloud() and
//   getOccupiedVoxelCenters().
// The octree's voxel lattice is anchored at the FIRST finite point (its bounding box starts as p0 -+ resolution/2 and only ever
// grows by whole multiples of the voxel size), so the set of occupied voxel centres is
//     { (floor((p - a) / r) + 0.5) r + a },  a = p0 - r/2  (float64, as PCL keeps min_x_/resolution_)
// independent of the insertion order.  Here: one transform kernel over all keyframes, 3 x 21-bit voxel keys, a radix sort of the
// keys (cub::DeviceRadixSort — 10^7..10^8 keys, far beyond the cluster sort's reach), unique, centres.  The result is the same SET
// of centres as PCL's (float rounding of the centre aside); PCL emits them in octree traversal order, here they come out in
// ascending (z, y, x) key order — the consumers (rviz, save_map) do not depend on the order.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include "engine.cuh"

namespace b2r {

struct MapKf {            // one keyframe snapshot on the device
  const float* pts;       // records (device)
  long long first;        // index of its first point in the concatenated cloud
  int n;
  int pad;
  float T[12];            // rows 0..2 of pose.cast<float>()
};

struct MapGeom {
  double a[3];            // lattice anchor p0 - r/2
  double r;
  long long kmin[3];      // smallest voxel coordinate per axis
  int has_anchor;
  int overflow;           // extent / resolution does not fit 21 bits per axis
};

// dst = pose * src (Eigen Matrix4f * Vector4f, column accumulation ((m0 x + m1 y) + m2 z) + m3, w = 1), other fields copied
__global__ void k_map_transform(const MapKf* __restrict__ kfs, int n_kf, int stride_f, long long total, float* __restrict__ out) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_kf - 1;  // keyframe of point g: last one with first <= g
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (kfs[mid].first <= g) lo = mid; else hi = mid - 1; }
    const MapKf& K = kfs[lo];
    const float* p = K.pts + (size_t)(g - K.first) * stride_f;
    float* o = out + (size_t)g * stride_f;
    const float x = p[0], y = p[1], z = p[2];
    for (int k = 3; k < stride_f; k++) o[k] = p[k];
    o[0] = xform_row(K.T[0], K.T[1], K.T[2], K.T[3], x, y, z);
    o[1] = xform_row(K.T[4], K.T[5], K.T[6], K.T[7], x, y, z);
    o[2] = xform_row(K.T[8], K.T[9], K.T[10], K.T[11], x, y, z);
    if (stride_f >= 4) o[3] = 1.0f;
  }
}

// first finite point (lowest index) and the bounding box of the finite points (ordered ints)
__global__ void k_map_first_and_bbox(const float* __restrict__ cloud, int stride_f, long long total, unsigned long long* first_idx, int* mm) {
  int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
  int mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  unsigned long long fi = 0xffffffffffffffffull;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const float* p = cloud + (size_t)g * stride_f;
    const float x = p[0], y = p[1], z = p[2];
    if (!finite3(x, y, z)) continue;
    if ((unsigned long long)g < fi) fi = (unsigned long long)g;
    const int ox = f2ord(x), oy = f2ord(y), oz = f2ord(z);
    mn[0] = min(mn[0], ox); mn[1] = min(mn[1], oy); mn[2] = min(mn[2], oz);
    mx[0] = max(mx[0], ox); mx[1] = max(mx[1], oy); mx[2] = max(mx[2], oz);
  }
#pragma unroll
  for (int d = 0; d < 3; d++) { mn[d] = __reduce_min_sync(0xffffffffu, mn[d]); mx[d] = __reduce_max_sync(0xffffffffu, mx[d]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, fi, o); fi = t < fi ? t : fi; }
  if ((threadIdx.x & 31) == 0) {
    if (fi != 0xffffffffffffffffull) atomicMin(first_idx, fi);
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (mn[d] != 0x7fffffff) atomicMin(&mm[d], mn[d]);
      if (mx[d] != (int)0x80000000) atomicMax(&mm[3 + d], mx[d]);
    }
  }
}

__global__ void k_map_geom(const float* __restrict__ cloud, int stride_f, const unsigned long long* first_idx, const int* mm, double r, MapGeom* g) {
  MapGeom G;
  G.r = r; G.has_anchor = 0; G.overflow = 0;
  for (int d = 0; d < 3; d++) { G.a[d] = 0; G.kmin[d] = 0; }
  if (*first_idx != 0xffffffffffffffffull) {
    const float* p = cloud + (size_t)(*first_idx) * stride_f;
    G.has_anchor = 1;
    for (int d = 0; d < 3; d++) {
      G.a[d] = (double)p[d] - r / 2;  // OctreePointCloud::adoptBoundingBoxToPoint, first point: min = point - resolution / 2
      const double lo = floor(((double)ord2f(mm[d]) - G.a[d]) / r), hi = floor(((double)ord2f(mm[3 + d]) - G.a[d]) / r);
      G.kmin[d] = (long long)lo;
      if (!(hi - lo < 2097151.0)) G.overflow = 1;
    }
  }
  *g = G;
}

__global__ void k_map_keys(const float* __restrict__ cloud, int stride_f, long long total, const MapGeom* __restrict__ gp, unsigned long long* keys) {
  const MapGeom G = *gp;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const float* p = cloud + (size_t)g * stride_f;
    const float x = p[0], y = p[1], z = p[2];
    unsigned long long key = 0xffffffffffffffffull;  // non-finite points are not inserted (OctreePointCloud::addPointsFromInputCloud)
    if (finite3(x, y, z) && G.has_anchor && !G.overflow) {
      // genOctreeKeyforPoint: key = (unsigned)((point - min) / resolution) in float64; min = a - (whole voxels)
      const unsigned long long kx = (unsigned long long)((long long)floor(((double)x - G.a[0]) / G.r) - G.kmin[0]);
      const unsigned long long ky = (unsigned long long)((long long)floor(((double)y - G.a[1]) / G.r) - G.kmin[1]);
      const unsigned long long kz = (unsigned long long)((long long)floor(((double)z - G.a[2]) / G.r) - G.kmin[2]);
      key = (kz << 42) | (ky << 21) | kx;
    }
    keys[g] = key;
  }
}

// sorted keys -> one PointXYZI-shaped record per distinct key: the voxel centre (genLeafNodeCenterFromOctreeKey), data[3] = 1,
// intensity 0; `count` is bumped with one atomic per block (the slot of a head = number of heads before it)
__global__ void k_map_heads(const unsigned long long* __restrict__ keys, long long total, int* flags) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[g];
    flags[g] = (k != 0xffffffffffffffffull && (g == 0 || keys[g - 1] != k)) ? 1 : 0;
  }
}
__global__ void k_map_centers(const unsigned long long* __restrict__ keys, const int* __restrict__ flags, const int* __restrict__ slots, long long total,
                              const MapGeom* __restrict__ gp, int stride_f, float* __restrict__ out, int* n_out) {
  const MapGeom G = *gp;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    if (g == total - 1) *n_out = slots[g] + flags[g];
    if (!flags[g]) continue;
    const unsigned long long k = keys[g];
    const long long kx = (long long)(k & 0x1fffffull) + G.kmin[0], ky = (long long)((k >> 21) & 0x1fffffull) + G.kmin[1], kz = (long long)(k >> 42) + G.kmin[2];
    float* o = out + (size_t)slots[g] * stride_f;
    o[0] = (float)(((double)kx + 0.5) * G.r + G.a[0]);
    o[1] = (float)(((double)ky + 0.5) * G.r + G.a[1]);
    o[2] = (float)(((double)kz + 0.5) * G.r + G.a[2]);
    if (stride_f >= 4) o[3] = 1.0f;
    for (int q = 4; q < stride_f; q++) o[q] = 0.f;
  }
}

}  // namespace b2r
