// voxelgrid.cuh — voxel-grid downsample, the companion kernel of the prefilter.
//
// Re-creates pcl::VoxelGrid<pcl::PointXYZI>::filter with setLeafSize(r,r,r), downsample_all_data = true, as
// configured at /root/reference/apps/prefiltering_nodelet.cpp:54-58 and called at :138-149 (and, optionally, at
// apps/scan_matching_odometry_nodelet.cpp:86-90,147-157):
//   k_vg_params    <- getMinMax3D + min_b/div_b/divb_mul + the int32 overflow guard ("Leaf size is too small ...")
//   k_vg_keys      <- idx = (floor(p * inv_leaf) - min_b) . divb_mul          (float32 multiply, int32 index)
//   radix sort     <- std::sort of (idx, point index): stable LSD radix sort => ascending point index inside a voxel
//   k_vg_centroid  <- one thread per voxel, sequential float32 sums of x,y,z,intensity in that order, / count
// Output order = ascending voxel key (as PCL).  Keys, counts and ordering are bit-exact against the oracle; centroids too,
// because the within-voxel summation order is pinned (PCL leaves it implementation-defined).
// The sort is the CUDA toolkit's cub::DeviceRadixSort (a library primitive, like a plain cuBLAS call); every other stage
// is hand-written.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <climits>
#include <vector>
#include "engine.cuh"

namespace b2r {

struct VgGeom {
  float inv_leaf;
  int min_b[3], div_b[3];
  int mul1, mul2;
  int overflow;  // dx*dy*dz > INT_MAX
  int empty;
};

struct VoxelWork {
  DevBuf<float> in, out;
  DevBuf<int> keys_a, keys_b, vals_a, vals_b, flags, slots, okeys, ocounts;
  DevBuf<char> tmp;
  int* mm = nullptr;
  VgGeom* geom = nullptr;
  VgGeom* h_geom = nullptr;  // pinned
  int* h_total = nullptr;    // pinned
  Telemetry* tel = nullptr;
  void release() {
    in.release(); out.release(); keys_a.release(); keys_b.release(); vals_a.release(); vals_b.release(); flags.release(); slots.release();
    okeys.release(); ocounts.release(); tmp.release();
    if (mm) cudaFree(mm);
    if (geom) cudaFree(geom);
    if (h_geom) cudaFreeHost(h_geom);
    if (h_total) cudaFreeHost(h_total);
    mm = nullptr; geom = nullptr; h_geom = nullptr; h_total = nullptr;
  }
};

__global__ void k_vg_params(const int* mm, VgGeom* g, int n, float leaf) {
  VgGeom G;
  G.inv_leaf = 1.0f / leaf;
  G.overflow = 0;
  G.empty = 0;
  G.mul1 = G.mul2 = 1;
  for (int d = 0; d < 3; d++) { G.min_b[d] = 0; G.div_b[d] = 1; }
  if (n <= 0 || mm[0] == 0x7fffffff) { G.empty = 1; *g = G; return; }
  long long vol = 1;
  bool huge = false;
  for (int d = 0; d < 3; d++) {
    float mn = ord2f(mm[d]), mx = ord2f(mm[3 + d]);
    float ext = fmul(fsub(mx, mn), G.inv_leaf);
    if (!(ext < 4.0e9f)) { huge = true; continue; }
    long long dd = (long long)ext + 1;  // static_cast<int64>((max - min) * inv_leaf) + 1
    vol *= dd;
    if (vol > (long long)INT_MAX) huge = true;
  }
  if (huge) { G.overflow = 1; *g = G; return; }
  for (int d = 0; d < 3; d++) {
    float mn = ord2f(mm[d]), mx = ord2f(mm[3 + d]);
    G.min_b[d] = (int)floorf(fmul(mn, G.inv_leaf));
    int max_b = (int)floorf(fmul(mx, G.inv_leaf));
    G.div_b[d] = max_b - G.min_b[d] + 1;
  }
  G.mul1 = G.div_b[0];
  G.mul2 = G.div_b[0] * G.div_b[1];
  *g = G;
}

__global__ void k_vg_keys(const float* __restrict__ raw, int stride_f, int n, const VgGeom* __restrict__ gp, int* keys, int* vals) {
  const VgGeom G = *gp;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = raw + (size_t)i * stride_f;
  const float x = p[0], y = p[1], z = p[2];
  int key = INT_MAX;  // non-finite points sort last and are dropped
  if (finite3(x, y, z) && !G.overflow) {
    int i0 = (int)fsub(floorf(fmul(x, G.inv_leaf)), (float)G.min_b[0]);
    int i1 = (int)fsub(floorf(fmul(y, G.inv_leaf)), (float)G.min_b[1]);
    int i2 = (int)fsub(floorf(fmul(z, G.inv_leaf)), (float)G.min_b[2]);
    key = i0 + i1 * G.mul1 + i2 * G.mul2;
  }
  keys[i] = key;
  vals[i] = i;
}

__global__ void k_vg_heads(const int* __restrict__ keys, int n, int* flags) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int k = keys[s];
  flags[s] = (k != INT_MAX && (s == 0 || keys[s - 1] != k)) ? 1 : 0;
}

__global__ void k_vg_centroid(const float* __restrict__ raw, int stride_f, int n, const int* __restrict__ keys, const int* __restrict__ vals,
                              const int* __restrict__ flags, const int* __restrict__ slots, float* out, int out_stride_f, int* okeys,
                              int* ocounts, int* total) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (s == n - 1) *total = slots[s] + flags[s];
  if (!flags[s]) return;
  const int k = keys[s];
  const bool has_i = stride_f >= 5;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  int e = s;
  while (e < n && keys[e] == k) {
    const float* p = raw + (size_t)vals[e] * stride_f;
    sx = fadd(sx, p[0]); sy = fadd(sy, p[1]); sz = fadd(sz, p[2]);
    si = fadd(si, has_i ? p[4] : 0.f);
    e++;
  }
  const float cnt = (float)(e - s);
  const int slot = slots[s];
  float* o = out + (size_t)slot * out_stride_f;
  o[0] = __fdiv_rn(sx, cnt); o[1] = __fdiv_rn(sy, cnt); o[2] = __fdiv_rn(sz, cnt);
  if (out_stride_f >= 4) o[3] = 1.0f;
  if (out_stride_f >= 5) o[4] = __fdiv_rn(si, cnt);
  for (int j = 5; j < out_stride_f; j++) o[j] = 0.f;
  okeys[slot] = k;
  ocounts[slot] = e - s;
}

inline int voxelgrid_filter(VoxelWork& W, cudaStream_t st, const void* in, size_t n, size_t stride_bytes, float leaf, void* out,
                            size_t* n_out, int32_t* out_keys, int32_t* out_counts) {
  *n_out = 0;
  if (n == 0) return B2R_OK;
  if (n > (size_t)0x3fffffff) return fail(B2R_EINVAL, "too many points");
  const int sf = (int)(stride_bytes / 4);
  const int N = (int)n;
  if (!W.mm) {
    B2R_CUDA(cudaMalloc(&W.mm, 8 * sizeof(int)));
    B2R_CUDA(cudaMalloc(&W.geom, sizeof(VgGeom)));
    B2R_CUDA(cudaMallocHost(&W.h_geom, sizeof(VgGeom)));
    B2R_CUDA(cudaMallocHost(&W.h_total, sizeof(int)));
  }
  B2R_CUDA(W.in.reserve(n * sf));
  B2R_CUDA(W.out.reserve(n * sf));
  B2R_CUDA(W.keys_a.reserve(n)); B2R_CUDA(W.keys_b.reserve(n)); B2R_CUDA(W.vals_a.reserve(n)); B2R_CUDA(W.vals_b.reserve(n));
  B2R_CUDA(W.flags.reserve(n)); B2R_CUDA(W.slots.reserve(n)); B2R_CUDA(W.okeys.reserve(n)); B2R_CUDA(W.ocounts.reserve(n));
  size_t tmp_sort = 0, tmp_scan = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, W.keys_a.p, W.keys_b.p, W.vals_a.p, W.vals_b.p, N, 0, 32, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, W.flags.p, W.slots.p, N, st);
  B2R_CUDA(W.tmp.reserve(std::max(tmp_sort, tmp_scan) + 256));
  B2R_CUDA(cudaMemcpyAsync(W.in.p, in, n * stride_bytes, cudaMemcpyHostToDevice, st));
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (W.tel) W.tel->h2d += n * stride_bytes;
  TEL_BEGIN(W.tel, st);
  k_grid_reset<<<1, 32, 0, st>>>(W.mm);
  k_bbox<<<nb > 1184 ? 1184 : nb, 256, 0, st>>>(W.in.p, sf, N, W.mm);
  k_vg_params<<<1, 1, 0, st>>>(W.mm, W.geom, N, leaf);
  B2R_CUDA(cudaMemcpyAsync(W.h_geom, W.geom, sizeof(VgGeom), cudaMemcpyDeviceToHost, st));
  k_vg_keys<<<nb, 256, 0, st>>>(W.in.p, sf, N, W.geom, W.keys_a.p, W.vals_a.p);
  size_t tb = W.tmp.cap;
  cub::DeviceRadixSort::SortPairs(W.tmp.p, tb, W.keys_a.p, W.keys_b.p, W.vals_a.p, W.vals_b.p, N, 0, 32, st);
  k_vg_heads<<<nb, 256, 0, st>>>(W.keys_b.p, N, W.flags.p);
  tb = W.tmp.cap;
  cub::DeviceScan::ExclusiveSum(W.tmp.p, tb, W.flags.p, W.slots.p, N, st);
  k_vg_centroid<<<nb, 256, 0, st>>>(W.in.p, sf, N, W.keys_b.p, W.vals_b.p, W.flags.p, W.slots.p, W.out.p, sf, W.okeys.p, W.ocounts.p,
                                    (int*)(W.mm + 6));
  TEL_END(W.tel, KC_VOXELGRID, 14, st);
  B2R_CUDA(cudaGetLastError());
  B2R_CUDA(cudaMemcpyAsync(W.h_total, W.mm + 6, sizeof(int), cudaMemcpyDeviceToHost, st));
  B2R_CUDA(cudaStreamSynchronize(st));
  if (W.h_geom->overflow) {  // PCL: warn and pass the input through unchanged
    std::memcpy(out, in, n * stride_bytes);
    *n_out = n;
    if (out_keys) for (size_t i = 0; i < n; i++) out_keys[i] = -1;
    if (out_counts) for (size_t i = 0; i < n; i++) out_counts[i] = 1;
    return 1;
  }
  const size_t m = (size_t)*W.h_total;
  if (m) {
    B2R_CUDA(cudaMemcpyAsync(out, W.out.p, m * stride_bytes, cudaMemcpyDeviceToHost, st));
    if (out_keys) B2R_CUDA(cudaMemcpyAsync(out_keys, W.okeys.p, m * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (out_counts) B2R_CUDA(cudaMemcpyAsync(out_counts, W.ocounts.p, m * sizeof(int), cudaMemcpyDeviceToHost, st));
    B2R_CUDA(cudaStreamSynchronize(st));
  }
  if (W.tel) W.tel->d2h += m * stride_bytes;
  *n_out = m;
  return B2R_OK;
}

}  // namespace b2r
