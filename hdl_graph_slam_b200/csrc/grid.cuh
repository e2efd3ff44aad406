// grid.cuh — device-side construction of the uniform search grid (counting sort by cell, deterministic order).
//
// Plays the role of the kd-tree builds the reference pays at every setInputSource/setInputTarget
// (pcl::search::KdTree::setInputCloud inside fast_gicp; pcl::Registration::initCompute's FLANN tree;
// call sites apps/scan_matching_odometry_nodelet.cpp:172,177,246, include/hdl_graph_slam/loop_detector.hpp:122,136).
// Everything is computed on the device (bbox, cell size, dims) so a build never synchronises with the host.
#pragma once
#include "common.cuh"

namespace b2r {

constexpr int kCellCap = 1 << 23;          // dense cell table capacity (32 MB of int32)
constexpr int kScanItems = 8;              // cells per thread in the scan
constexpr int kScanThreads = 256;
constexpr int kScanTile = kScanItems * kScanThreads;  // 2048 cells per block
constexpr int kScanBlocks = kCellCap / kScanTile;     // 4096
constexpr int kScanBPerThread = kScanBlocks / 1024;   // block sums handled per thread in phase B

struct GridBuffers {
  Grid* grid;        // device Grid
  int* mm;           // 6 ordered ints (min xyz, max xyz)
  int* counts;       // kCellCap (kept zeroed between builds)
  int* cell_start;   // kCellCap + 1
  int* cursor;       // kCellCap
  int* bsum;         // kScanBlocks
  int* cell_of;      // n (cell id per original point, -1 = non-finite)
  int* tmp_idx;      // n
  float4* sorted;    // n  (x, y, z, bits(original index)), ascending (cell, original index)
  int* pos_of;       // n  original index -> sorted position (-1 if dropped)
};

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

// choose the cell size (power of two, >= h_min) so that the dense table fits kCellCap
__global__ void k_grid_params(const int* mm, Grid* g, int n, float h_min) {
  Grid G;
  G.n = n;
  G.pad = 0;
  if (n <= 0 || mm[0] == 0x7fffffff) {
    G.ox = G.oy = G.oz = 0.f; G.h = h_min; G.inv_h = 1.f / h_min; G.nx = G.ny = G.nz = 1; G.ncell = 1; G.n_valid = 0;
    *g = G;
    return;
  }
  float mn[3] = {ord2f(mm[0]), ord2f(mm[1]), ord2f(mm[2])};
  float mx[3] = {ord2f(mm[3]), ord2f(mm[4]), ord2f(mm[5])};
  float h = h_min;
  for (int it = 0; it < 64; it++) {
    float inv = 1.f / h;
    float o[3];
    double cells = 1.0;
    int dims[3];
    bool ok = true;
    for (int d = 0; d < 3; d++) {
      o[d] = floorf(mn[d] * inv) * h;  // multiple of h (exact: power-of-two scaling)
      float ext = floorf((mx[d] - o[d]) * inv);
      if (!(ext < 4.0e6f)) { ok = false; break; }
      dims[d] = (int)ext + 1;
      cells *= (double)dims[d];
    }
    if (ok && cells <= (double)kCellCap) {
      G.ox = o[0]; G.oy = o[1]; G.oz = o[2];
      G.h = h; G.inv_h = inv;
      G.nx = dims[0]; G.ny = dims[1]; G.nz = dims[2];
      G.ncell = dims[0] * dims[1] * dims[2];
      G.n_valid = 0;
      *g = G;
      return;
    }
    h *= 2.f;
  }
  // unreachable for finite input; fall back to a single cell
  G.ox = mn[0]; G.oy = mn[1]; G.oz = mn[2]; G.h = 3.0e38f; G.inv_h = 0.f; G.nx = G.ny = G.nz = 1; G.ncell = 1; G.n_valid = 0;
  *g = G;
}

__global__ void k_count(const float* __restrict__ raw, int stride_f, int n, const Grid* __restrict__ gp, int* counts, int* cell_of) {
  const Grid g = *gp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = raw + (size_t)i * stride_f;
    float x = p[0], y = p[1], z = p[2];
    int c = -1;
    if (finite3(x, y, z)) {
      int cx = cell_coord(x, g.ox, g.inv_h, g.nx);
      int cy = cell_coord(y, g.oy, g.inv_h, g.ny);
      int cz = cell_coord(z, g.oz, g.inv_h, g.nz);
      c = (cz * g.ny + cy) * g.nx + cx;
      atomicAdd(&counts[c], 1);
    }
    cell_of[i] = c;
  }
}

// ---- three-phase exclusive scan over the first g->ncell cells --------------------------------------------------
__device__ __forceinline__ int block_excl_scan(int v, int* smem /* >= 32 ints */, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int nw = (blockDim.x + 31) >> 5;
    int w = lane < nw ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  int res = inc - v + smem[warp];
  total = smem[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_a(const int* __restrict__ counts, const Grid* __restrict__ gp, int* bsum) {
  __shared__ int sm[40];
  const int ncell = gp->ncell;
  const int base = blockIdx.x * kScanTile;
  if (base >= ncell) { if (threadIdx.x == 0) bsum[blockIdx.x] = 0; return; }
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    int c = base + threadIdx.x * kScanItems + k;
    if (c < ncell) s += counts[c];
  }
  int total;
  block_excl_scan(s, sm, total);
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_b(int* bsum, Grid* gp, int* cell_start) {
  __shared__ int sm[40];
  int v[kScanBPerThread];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanBPerThread; k++) { v[k] = bsum[threadIdx.x * kScanBPerThread + k]; s += v[k]; }
  int total;
  int ex = block_excl_scan(s, sm, total);
#pragma unroll
  for (int k = 0; k < kScanBPerThread; k++) { bsum[threadIdx.x * kScanBPerThread + k] = ex; ex += v[k]; }
  if (threadIdx.x == 0) {
    gp->n_valid = total;
    cell_start[gp->ncell] = total;
  }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_c(int* counts, const Grid* __restrict__ gp, const int* __restrict__ bsum,
                                                          int* cell_start, int* cursor) {
  __shared__ int sm[40];
  const int ncell = gp->ncell;
  const int base = blockIdx.x * kScanTile;
  if (base >= ncell) return;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    int c = base + threadIdx.x * kScanItems + k;
    v[k] = (c < ncell) ? counts[c] : 0;
    s += v[k];
  }
  int total;
  int ex = block_excl_scan(s, sm, total) + bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    int c = base + threadIdx.x * kScanItems + k;
    if (c < ncell) {
      cell_start[c] = ex;
      cursor[c] = ex;
      counts[c] = 0;  // leave the table clean for the next build
      ex += v[k];
    }
  }
}

__global__ void k_scatter(int n, const int* __restrict__ cell_of, int* cursor, int* tmp_idx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int c = cell_of[i];
    if (c >= 0) tmp_idx[atomicAdd(&cursor[c], 1)] = i;
  }
}

// canonical (deterministic) order inside each cell: ascending original index
__global__ void k_canon(const float* __restrict__ raw, int stride_f, int n, const Grid* __restrict__ gp, const int* __restrict__ cell_of,
                        const int* __restrict__ cell_start, const int* __restrict__ tmp_idx, float4* sorted, int* pos_of) {
  const int nv = gp->n_valid;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    if (s >= nv) continue;
    int i = tmp_idx[s];
    int c = cell_of[i];
    int b = cell_start[c], e = cell_start[c + 1];
    int rank = 0;
    for (int j = b; j < e; j++) rank += (tmp_idx[j] < i) ? 1 : 0;
    const float* p = raw + (size_t)i * stride_f;
    sorted[b + rank] = make_float4(p[0], p[1], p[2], bits_idx(i));
    pos_of[i] = b + rank;
  }
}

__global__ void k_fill_i32(int* p, int n, int v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

// Enqueue a full grid build on `st`.  raw: device array of n records, stride_f floats each (x,y,z first).
inline void build_grid(const float* raw, int stride_f, int n, float h_min, const GridBuffers& B, cudaStream_t st) {
  k_grid_reset<<<1, 32, 0, st>>>(B.mm);
  int nb = n > 0 ? (n + 255) / 256 : 1;
  if (nb > 1184) nb = 1184;
  if (n > 0) k_bbox<<<nb, 256, 0, st>>>(raw, stride_f, n, B.mm);
  k_grid_params<<<1, 1, 0, st>>>(B.mm, B.grid, n, h_min);
  if (n > 0) {
    k_fill_i32<<<nb, 256, 0, st>>>(B.pos_of, n, -1);
    k_count<<<nb, 256, 0, st>>>(raw, stride_f, n, B.grid, B.counts, B.cell_of);
  }
  k_scan_a<<<kScanBlocks, kScanThreads, 0, st>>>(B.counts, B.grid, B.bsum);
  k_scan_b<<<1, 1024, 0, st>>>(B.bsum, B.grid, B.cell_start);
  k_scan_c<<<kScanBlocks, kScanThreads, 0, st>>>(B.counts, B.grid, B.bsum, B.cell_start, B.cursor);
  if (n > 0) {
    k_scatter<<<nb, 256, 0, st>>>(n, B.cell_of, B.cursor, B.tmp_idx);
    k_canon<<<nb, 256, 0, st>>>(raw, stride_f, n, B.grid, B.cell_of, B.cell_start, B.tmp_idx, B.sorted, B.pos_of);
  }
}

}  // namespace b2r
