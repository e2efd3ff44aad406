// engine.cuh — handle state of libb200reg: per-cloud device buffers, workspaces, stream.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include "../../include/b200reg.h"
#include "common.cuh"
#include "grid.cuh"

namespace b2r {

extern thread_local std::string g_last_error;
inline int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define B2R_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      return b2r::fail(B2R_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
    }                                                                                                    \
  } while (0)

// growable device buffer
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    cap = (e == cudaSuccess) ? want : 0;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct NdtVoxelMap;  // ndt.cuh

// One point cloud resident on the device with everything derived from it.
struct Cloud {
  size_t n = 0;
  int stride_f = 4;
  const void* host_ptr = nullptr;  // identity of the host buffer last uploaded (promotion / reuse detection)
  DevBuf<float> raw;               // uploaded records
  const float* raw_view = nullptr; // == raw.p, or a caller-owned device pointer (set_*_device)
  // search grid
  Grid* grid = nullptr;
  DevBuf<int> cell_start;          // kCellCap + 1
  DevBuf<float4> sorted;
  DevBuf<int> pos_of;
  bool grid_ready = false;
  // GICP
  DevBuf<double> cov;              // sorted order, 6 per point
  bool cov_ready = false;
  // NDT (target only)
  NdtVoxelMap* ndt = nullptr;
  bool ndt_ready = false;
  void invalidate() { grid_ready = cov_ready = ndt_ready = false; }
};

struct Scratch {  // build scratch shared by all builds of a handle (stream-ordered)
  int* mm = nullptr;
  int* counts = nullptr;
  int* cursor = nullptr;
  int* bsum = nullptr;
  DevBuf<int> cell_of, tmp_idx;
};

}  // namespace b2r
