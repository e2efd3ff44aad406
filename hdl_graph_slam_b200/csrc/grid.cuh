// grid.cuh — small device-side helpers shared by every structure build (BVH, NDT voxel map): the cloud's bounding box as
// ordered ints (so that float min / max can use integer atomics) and a fill kernel.  Everything is computed on the device, so a
// build never synchronises with the host.  (The dense uniform grid of the first design is gone: the searches run on the
// implicit BVH of bvh.cuh, the NDT voxel map is sparse — ndt.cuh.)
#pragma once
#include "common.cuh"

namespace b2r {

__global__ void k_grid_reset(int* mm) {
  if (threadIdx.x < 3) mm[threadIdx.x] = 0x7fffffff;
  else if (threadIdx.x < 6) mm[threadIdx.x] = (int)0x80000000;
}

__global__ void k_bbox(const float* __restrict__ raw, int stride_f, int n, int* mm) {
  int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
  int mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = raw + (size_t)i * stride_f;
    float x = p[0], y = p[1], z = p[2];
    if (!finite3(x, y, z)) continue;
    int ox = f2ord(x), oy = f2ord(y), oz = f2ord(z);
    mn[0] = min(mn[0], ox); mn[1] = min(mn[1], oy); mn[2] = min(mn[2], oz);
    mx[0] = max(mx[0], ox); mx[1] = max(mx[1], oy); mx[2] = max(mx[2], oz);
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (mn[d] != 0x7fffffff) atomicMin(&mm[d], mn[d]);
      if (mx[d] != (int)0x80000000) atomicMax(&mm[3 + d], mx[d]);
    }
  }
}

__global__ void k_fill_i32(int* p, int n, int v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace b2r
