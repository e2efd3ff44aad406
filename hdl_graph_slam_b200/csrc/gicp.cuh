// gicp.cuh — GICP kernels: k-NN covariances, correspondences, linearisation (+ fused LM trial cost), final trial cost.
//
// Re-creates, B200-first, the arithmetic of fast_gicp::FastGICP that the reference selects at
// /root/reference/src/hdl_graph_slam/registrations.cpp:27-36 and runs from apps/scan_matching_odometry_nodelet.cpp:177,210
// and include/hdl_graph_slam/loop_detector.hpp:136,143:
//   k_knn_cov_reg<20>  <- FastGICP::calculate_covariances (k = 20 exact NN in registers, PLANE regularisation); k_knn_cov: any k
//   k_gicp_correspond  <- FastGICP::update_correspondences (exact seeded 1-NN, float32, kNnCopies lanes per query)
//   k_gicp_accumulate  <- FastGICP::linearize (float64; M_i kept) fused with FastGICP::compute_error of the previous set (trial cost)
//   k_gicp_error       <- FastGICP::compute_error
// All per-cloud arrays live in the BVH's sorted order (Hilbert key, then original index), so consecutive threads hold spatial
// neighbours; block partials are combined in a fixed order => bitwise reproducible results.
#pragma once
#include "common.cuh"
#include "bvh.cuh"
#include "linalg.cuh"

namespace b2r {

constexpr int kKnnThreads = 128;
constexpr int kLinThreads = 128;
#ifndef B2R_ACC_THREADS
#define B2R_ACC_THREADS 256
#endif
constexpr int kAccThreads = B2R_ACC_THREADS;  // k_gicp_accumulate: ~126 registers per thread; 256 rows of partials for the final reduction
constexpr int kAcc = 29;  // 21 (upper H) + 6 (b) + 1 (cost) + 1 (trial cost at this pose with the PREVIOUS correspondences)

// ---------------------------------------------------------------- k-NN covariance
// per-lane sorted top-k of packed 64-bit keys in shared memory, slot-major ([k][blockDim]); the k-th key is cached in a
// register so the all-pairs tile loop costs one 64-bit compare per candidate.
struct KnnList {
  static constexpr int kTileLanes = 3;   // coop mode only when 1-2 lanes want the leaf (each pick costs an insertion)
  static constexpr int kTileUnroll = 2;  // the insertion loop is big: keep the instruction footprint small (i-cache)
  static constexpr bool kTwoPhase = true;
  unsigned long long* key;
  int k, cnt, stride;
  unsigned long long wkey;  // key[k-1] once the list is full, else kKeyInf
#ifdef B2R_KNN_PROFILE
  int n_test = 0, n_ins = 0, n_shift = 0, n_tile = 0, n_coop = 0, n_try = 0, n_obb = 0;
#endif
  __device__ __forceinline__ float worst() const { return nn_key_d2(wkey); }
  __device__ __forceinline__ float limit() const { return INFINITY; }
  __device__ __forceinline__ void visit(float d2, int idx, int) {
    const unsigned long long kq = nn_key(d2, idx);
#ifdef B2R_KNN_PROFILE
    n_test++;
#endif
    if (kq >= wkey) return;
#ifdef B2R_KNN_PROFILE
    n_ins++;
#endif
    int j = (cnt < k) ? cnt++ : k - 1;
    while (j > 0) {
      const unsigned long long prev = key[(j - 1) * stride];
      if (prev <= kq) break;
      key[j * stride] = prev;
      j--;
#ifdef B2R_KNN_PROFILE
      n_shift++;
#endif
    }
    key[j * stride] = kq;
    if (cnt == k) wkey = key[(k - 1) * stride];
  }
};

// mean / covariance over the neighbours in ascending (d2, index) order (float64), PLANE regularisation, store 6 doubles.
// pt_at(j) returns the j-th neighbour's coordinates.
template <class PtAt>
__device__ __forceinline__ void knn_cov_store(int kk, PtAt pt_at, double* __restrict__ o) {
  double mx = 0, my = 0, mz = 0;
  for (int j = 0; j < kk; j++) {
    const float3 p = pt_at(j);
    mx += (double)p.x; my += (double)p.y; mz += (double)p.z;
  }
  const double inv = 1.0 / (double)kk;
  mx *= inv; my *= inv; mz *= inv;
  double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < kk; j++) {
    const float3 p = pt_at(j);
    double vx = (double)p.x - mx, vy = (double)p.y - my, vz = (double)p.z - mz;
    c[0] += vx * vx; c[1] += vx * vy; c[2] += vx * vz;
    c[4] += vy * vy; c[5] += vy * vz; c[8] += vz * vz;
  }
  c[0] *= inv; c[1] *= inv; c[2] *= inv; c[4] *= inv; c[5] *= inv; c[8] *= inv;
  c[3] = c[1]; c[6] = c[2]; c[7] = c[5];
  // PLANE regularisation: singular values (descending) replaced by (1, 1, 1e-3).  With orthonormal eigenvectors
  // v0 (smallest), v1, v2:  1*v2 v2^T + 1*v1 v1^T + 1e-3*v0 v0^T  ==  I - (1 - 1e-3) v0 v0^T : only the normal is needed.
  double n[3];
  sym_min_eigvec3(c, n);
  const double w = 1.0 - 1e-3;
  o[0] = 1.0 - w * n[0] * n[0];
  o[1] = -w * n[0] * n[1];
  o[2] = -w * n[0] * n[2];
  o[3] = 1.0 - w * n[1] * n[1];
  o[4] = -w * n[1] * n[2];
  o[5] = 1.0 - w * n[2] * n[2];
}

// one warp per leaf (4 leaves per block): its 32 points are the queries; per-lane top-k lists in shared memory
__global__ void __launch_bounds__(kKnnThreads, 4) k_knn_cov(Bvh b, int k, const float* __restrict__ raw, int stride_f, double* __restrict__ cov,
                                                            long long* prof) {
  extern __shared__ unsigned long long knn_keys[];  // [k][blockDim]
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int leaf = s >> 5;
  if (leaf >= b.nleaf) return;  // whole warps only (blockDim is a multiple of 32)
  const float4 q = b.sp[s];
  const bool active = idx_bits(q.w) != kPadIdx;
  KnnList L;
  L.key = knn_keys + threadIdx.x;
  L.k = k;
  L.cnt = 0;
  L.stride = blockDim.x;
  L.wkey = kKeyInf;
#ifdef B2R_KNN_PROFILE
  const long long t0 = clock64();
#endif
  bvh_group_search(b, q.x, q.y, q.z, active, L, leaf);
#ifdef B2R_KNN_PROFILE
  {
    const long long t1 = clock64();
    const unsigned F = 0xffffffffu;
    int v[6] = {L.n_test, L.n_ins, L.n_shift, L.n_ins, L.n_shift, 0};
    for (int o = 16; o > 0; o >>= 1) {
      v[0] += __shfl_xor_sync(F, v[0], o); v[1] += __shfl_xor_sync(F, v[1], o); v[2] += __shfl_xor_sync(F, v[2], o);
      v[3] = max(v[3], __shfl_xor_sync(F, v[3], o)); v[4] = max(v[4], __shfl_xor_sync(F, v[4], o));
    }
    const int ntile = __shfl_sync(F, L.n_tile, 0), ncoop = __shfl_sync(F, L.n_coop, 0);
    if ((threadIdx.x & 31) == 0) {
      long long* o = prof + (size_t)leaf * 8;
      o[0] = t1 - t0; o[1] = ntile; o[2] = ncoop; o[3] = v[0]; o[4] = v[1]; o[5] = v[2]; o[6] = v[3]; o[7] = v[4];
    }
  }
#endif
  if (!active) return;
  const int stride = L.stride;
  const unsigned long long* keyp = L.key;
  knn_cov_store(L.cnt, [=](int j) {
    const int idx = (int)(unsigned int)(keyp[j * stride] & 0xffffffffull);
    const float* p = raw + (size_t)idx * stride_f;
    return make_float3(p[0], p[1], p[2]);
  }, cov + (size_t)s * 6);
}

// ---- register-resident lists (K fixed at compile time).  profiles/r01_g: with the lists in shared memory a warp spends ~400
// cycles per candidate step, the latency of a data-dependent LDS -> compare -> STS chain (max over lanes of the shift count);
// in registers every slot is recomputed branch-free from the OLD neighbours (new[j] = old[j-1] > q ? old[j-1] : old[j] > q ? q :
// old[j]), so the K slot updates are independent instructions.  Same candidates, same (d2, idx) order => identical lists.
template <int K>
struct KnnRegs {
  static constexpr int kTileLanes = 3;
  static constexpr int kTileUnroll = 1;
  static constexpr bool kTwoPhase = true;
  unsigned long long key[K];
#ifdef B2R_KNN_PROFILE
  int n_test = 0, n_ins = 0, n_shift = 0, n_tile = 0, n_coop = 0, n_try = 0, n_obb = 0;
#endif
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int j = 0; j < K; j++) key[j] = kKeyInf;
  }
  __device__ __forceinline__ float worst() const { return nn_key_d2(key[K - 1]); }
  __device__ __forceinline__ float limit() const { return INFINITY; }
  __device__ __forceinline__ void visit(float d2, int idx, int) {
    const unsigned long long kq = nn_key(d2, idx);
#ifdef B2R_KNN_PROFILE
    n_test++;
#endif
    if (kq >= key[K - 1]) return;  // also rejects padding (kq == kKeyInf)
#ifdef B2R_KNN_PROFILE
    n_ins++;
#endif
    bool pj = true;  // p_j = old key[j] > kq, monotone in j; p_{K-1} holds (tested above)
#pragma unroll
    for (int j = K - 1; j > 0; j--) {  // descending: key[j-1] is still the old value when slot j is rewritten
      const unsigned long long lo = key[j - 1];
      const bool pl = lo > kq;
      if (pj) key[j] = pl ? lo : kq;   // slot j changes only if p_j: takes its left neighbour, or the new key at the boundary
      pj = pl;
    }
    if (pj) key[0] = kq;
  }
  __device__ __forceinline__ int count() const {
    int c = 0;
#pragma unroll
    for (int j = 0; j < K; j++) c += key[j] != kKeyInf;
    return c;
  }
};

constexpr int kKnnRegK = 20;

template <int K>
__global__ void __launch_bounds__(kKnnThreads, 4) k_knn_cov_reg(Bvh b, const float* __restrict__ raw, int stride_f, double* __restrict__ cov,
                                                                long long* prof) {
  __shared__ int nbr[K][kKnnThreads];  // neighbour indices, ascending (d2, idx), for the covariance pass
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int leaf = s >> 5;
  if (leaf >= b.nleaf) return;  // whole warps only
  const float4 q = b.sp[s];
  const bool active = idx_bits(q.w) != kPadIdx;
  KnnRegs<K> L;
  L.reset();
#ifdef B2R_KNN_PROFILE
  const long long t0 = clock64();
#endif
  bvh_group_search(b, q.x, q.y, q.z, active, L, leaf);
#ifdef B2R_KNN_PROFILE
  {
    const long long t1 = clock64();
    const unsigned F = 0xffffffffu;
    int v[4] = {L.n_test, L.n_ins, L.n_ins, 0};
    for (int o = 16; o > 0; o >>= 1) {
      v[0] += __shfl_xor_sync(F, v[0], o); v[1] += __shfl_xor_sync(F, v[1], o);
      v[2] = max(v[2], __shfl_xor_sync(F, v[2], o));
    }
    const int ntile = __shfl_sync(F, L.n_tile, 0), ncoop = __shfl_sync(F, L.n_coop, 0);
    if ((threadIdx.x & 31) == 0) {
      long long* o = prof + (size_t)leaf * 8;
      o[0] = t1 - t0; o[1] = ntile; o[2] = ncoop; o[3] = v[0]; o[4] = v[1]; o[5] = 0; o[6] = v[2]; o[7] = 0;
    }
  }
#endif
  if (!active) return;
#pragma unroll
  for (int j = 0; j < K; j++) nbr[j][threadIdx.x] = (int)(unsigned int)(L.key[j] & 0xffffffffull);
  const int tx = threadIdx.x;
  knn_cov_store(L.count(), [&](int j) {
    const float* p = raw + (size_t)nbr[j][tx] * stride_f;
    return make_float3(p[0], p[1], p[2]);
  }, cov + (size_t)s * 6);
}

// ---------------------------------------------------------------- pose passed by value to the per-iteration kernels
struct PoseArg {
  double T[12];  // rows 0..2 of the 4x4 (row-major)
  float Tf[12];  // float cast of the same (fast_gicp: trans.cast<float>())
};

// deterministic block reduction of NV doubles per thread; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_reduce(double* v, double* smem /* NV * 32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double x = (lane < nw) ? smem[i * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      v[i] = x;
    }
  }
}

// last-block-done final reduction: partials [gridDim][NV] -> out[NV], summed in a fixed order.
// `out` may live in host-mapped pinned memory: the result then lands in host memory straight from the kernel together with a
// checksum word (out[NV+1]; out[NV] carries `extra` or is unused) and `flag` = `seq`, so the host can spin on it instead of paying a
// memcpy + stream sync.
template <int NV>
__device__ __forceinline__ void finish_partials(const double* v, double* partials, double* out, unsigned int* counter,
                                                unsigned long long* flag = nullptr, unsigned long long seq = 0,
                                                unsigned long long* extra = nullptr) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) partials[(size_t)blockIdx.x * NV + i] = v[i];
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // fixed-order parallel sum: warp w adds rows w, w+nw, ... (lane = value: one coalesced row per load, 8 loads in flight,
    // ld.cg: the rows were written by other SMs during this launch), then thread i adds the per-warp sums in warp order.
    // The order depends only on the launch geometry => run-to-run bitwise reproducible.
    __shared__ double fin[8 * NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = (blockDim.x >> 5) < 8 ? (int)(blockDim.x >> 5) : 8;
    const unsigned int nrow = gridDim.x;
    if (warp < nw) {
      for (int i = lane; i < NV; i += 32) {
        double s = 0.0;
        unsigned int r = warp;
        for (; r + 7 * nw < nrow; r += 8 * nw) {
          double t[8];
#pragma unroll
          for (int u = 0; u < 8; u++) t[u] = __ldcg(partials + (size_t)(r + u * nw) * NV + i);
#pragma unroll
          for (int u = 0; u < 8; u++) s += t[u];
        }
        for (; r < nrow; r += nw) s += __ldcg(partials + (size_t)r * NV + i);
        fin[warp * NV + i] = s;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NV; i += blockDim.x) {
      double s = 0.0;
      for (int w = 0; w < nw; w++) s += fin[w * NV + i];
      out[i] = s;
      fin[i] = s;  // column i of `fin` is read by this thread only: slot [0][i] is free again and carries the value to thread 0
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *counter = 0;
      unsigned long long x = seq;
      if (extra) {
        const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(extra);
        *extra = 0;  // device-side counter accumulated with atomics during this launch: read once, re-armed for the next launch
        reinterpret_cast<unsigned long long*>(out)[NV] = e;
        x ^= e;
      }
      if (flag) {
        // Publication without system-scope fences (each costs a PCIe round trip): the host accepts the result words only when
        // flag == seq AND the checksum word equals seq ^ xor(result words), so the order in which these stores land in host memory
        // does not matter (wait_host_result re-reads until the message is self-consistent).
        for (int i = 0; i < NV; i++) x ^= (unsigned long long)__double_as_longlong(fin[i]);
        reinterpret_cast<unsigned long long*>(out)[NV + 1] = x;
        *reinterpret_cast<volatile unsigned long long*>(flag) = seq;
      }
    }
  }
}

struct LinArgs {
  Bvh src;                     // source structure (queries = its sorted points)
  const double* scov;          // source covariances, sorted (6 per point)
  Bvh tgt;
  const double* tcov;
  double thr2;                 // max_correspondence_distance^2 (double, compared against (double)d2)
  float lim;                   // float >= thr2 (search range limit)
  int* corr;                   // [n_src original] target ORIGINAL index or -1
  int* cpos;                   // [sorted s] target sorted position or -1 (written)
  const int* cpos_prev;        // previous iteration's correspondences: search seeds, and the trial cost (compute_error) when fuse_error
  const double* mahal_prev;    // previous iteration's M_i (fuse_error)
  int fuse_error;              // 1: also accumulate sum e^T M_prev e over the previous correspondences into acc[28]
  float* d2;                   // [sorted s] NN squared distance
  double* mahal;               // [sorted s][6]
  double* partials;            // [blocks][kAcc]
  double* out;                 // [kAcc]
  unsigned int* counter;
  unsigned long long* flag;    // host-mapped completion flag (see finish_partials)
  unsigned long long seq;
  int use_seed;                // 1: cpos[] holds last iteration's correspondences -> seed the search bound
  const int* tgt_pos_of;       // first iteration of an align (no correspondences yet): the target point with the SAME original index
  int tgt_n;                   //   is the seed — consecutive scans of a spinning LiDAR share their firing order, so it is a near neighbour
  long long* prof;             // B2R_KNN_PROFILE builds only
};

// update_correspondences: exact 1-NN of every transformed source point (float32 search, few registers: it shares the SMs with the
// k-NN covariance kernel of the prefetched next scan).  Writes corr / cpos / d2; no reduction.
// kNnCopies lanes per query (bvh.cuh): a warp carries 32/kNnCopies queries -> that many times the warps, shorter chains per warp.
#ifndef B2R_NN_COPIES
#define B2R_NN_COPIES 4
#endif
constexpr int kNnCopies = B2R_NN_COPIES;
template <int C>
__global__ void __launch_bounds__(kLinThreads, 8) k_gicp_correspond(const __grid_constant__ LinArgs A, const __grid_constant__ PoseArg P) {
  constexpr int Q = 32 / C;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = (gt >> 5) * Q + (gt & (Q - 1));
  const bool writer = (gt & 31) < Q;
  float4 p = make_float4(0.f, 0.f, 0.f, bits_idx(kPadIdx));
  if (s < A.src.nleaf * kLeaf) p = A.src.sp[s];
  const bool is_point = idx_bits(p.w) != kPadIdx;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  Nn1 v;
  v.reset(A.lim);
  bool active = false;
  int sp0 = -1;
  if (is_point) {
    qx = xform_row(P.Tf[0], P.Tf[1], P.Tf[2], P.Tf[3], p.x, p.y, p.z);
    qy = xform_row(P.Tf[4], P.Tf[5], P.Tf[6], P.Tf[7], p.x, p.y, p.z);
    qz = xform_row(P.Tf[8], P.Tf[9], P.Tf[10], P.Tf[11], p.x, p.y, p.z);
    if (finite3(qx, qy, qz)) {
      active = true;
      // a seed is a REAL candidate (its exact distance and index enter the visitor like any other): a tight upper bound from the
      // last iteration's correspondent, or a heuristic one from the firing order on the first iteration — never a guess about the answer
      if (A.use_seed) sp0 = A.cpos_prev[s];
      else if (A.tgt_pos_of) { const int oi = idx_bits(p.w); if (oi < A.tgt_n) sp0 = A.tgt_pos_of[oi]; }
      {
        if (sp0 >= 0) {
          float4 t = A.tgt.sp[sp0];
          v.seed(dist2_f32(qx, qy, qz, t.x, t.y, t.z), idx_bits(t.w), sp0);
        }
      }
    }
  }
  // traversal hint: the leaf of the middle seeded lane's correspondent (Hilbert order: the group's answers sit around it).  A
  // firing-order seed is only trusted as a starting point when it is close (odometry: consecutive scans); for unrelated scan
  // pairs (loop closure) it is a far point and the centre-nearest leaf lookup is the better start — the bound it gives is kept.
  int hint = -1;
  {
    const bool good = sp0 >= 0 && (A.use_seed || v.bd2 < 1.0f);
    const unsigned hm = __ballot_sync(0xffffffffu, good);
    if (hm) hint = __shfl_sync(0xffffffffu, sp0, __fns(hm, 0, (__popc(hm) + 1) / 2)) >> 5;
  }
#ifdef B2R_KNN_PROFILE
  const long long t0 = clock64();
#endif
  bvh_group_search<C>(A.tgt, qx, qy, qz, active, v, -1, hint);  // all 32 lanes participate
#ifdef B2R_KNN_PROFILE
  if (A.prof && (threadIdx.x & 31) == 0) {
    long long* o = A.prof + (size_t)(gt >> 5) * 4;
    o[0] = clock64() - t0; o[1] = v.n_tile; o[2] = v.n_coop; o[3] = v.n_try;
  }
#endif
  if (is_point && writer) {
    const bool valid = active && (v.best_pos >= 0) && ((double)v.best_d2() < A.thr2);
    A.corr[idx_bits(p.w)] = valid ? v.best_idx() : -1;
    A.cpos[s] = valid ? v.best_pos : -1;
    A.d2[s] = v.best_d2();
  }
}

// linearize over the correspondences just written (float64), fused with the trial cost compute_error over the PREVIOUS set.
__global__ void __launch_bounds__(kAccThreads, 512 / kAccThreads) k_gicp_accumulate(const __grid_constant__ LinArgs A, const __grid_constant__ PoseArg P) {
  __shared__ double red[kAcc * 32];
  // launched with programmatic stream serialization right behind k_gicp_correspond: the blocks may be scheduled while the search
  // kernel drains, and wait here until it has completed and its cpos / corr / d2 writes are visible
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; i++) acc[i] = 0.0;
  float4 p = make_float4(0.f, 0.f, 0.f, bits_idx(kPadIdx));
  if (s < A.src.nleaf * kLeaf) p = A.src.sp[s];
  const bool is_point = idx_bits(p.w) != kPadIdx;
  if (A.fuse_error && is_point) {  // FastGICP::compute_error at this pose with the previous correspondences / mahalanobis
    const int tp = A.cpos_prev[s];
    if (tp >= 0) {
      const float4 tb = A.tgt.sp[tp];
      const double* m = A.mahal_prev + (size_t)s * 6;
      const double ax = (double)p.x, ay = (double)p.y, az = (double)p.z;
      const double ex = (double)tb.x - (P.T[0] * ax + P.T[1] * ay + P.T[2] * az + P.T[3]);
      const double ey = (double)tb.y - (P.T[4] * ax + P.T[5] * ay + P.T[6] * az + P.T[7]);
      const double ez = (double)tb.z - (P.T[8] * ax + P.T[9] * ay + P.T[10] * az + P.T[11]);
      const double Mex = m[0] * ex + m[1] * ey + m[2] * ez;
      const double Mey = m[1] * ex + m[3] * ey + m[4] * ez;
      const double Mez = m[2] * ex + m[4] * ey + m[5] * ez;
      acc[28] = ex * Mex + ey * Mey + ez * Mez;
    }
  }
  if (is_point) {
    const int best_pos = A.cpos[s];
    const bool valid = best_pos >= 0;
    if (valid) {
      const double* ca = A.scov + (size_t)s * 6;
      const double* cb = A.tcov + (size_t)best_pos * 6;
      const double CA[9] = {ca[0], ca[1], ca[2], ca[1], ca[3], ca[4], ca[2], ca[4], ca[5]};
      const double R[9] = {P.T[0], P.T[1], P.T[2], P.T[4], P.T[5], P.T[6], P.T[8], P.T[9], P.T[10]};
      double tmp[9], rcr[9], M[9];
      mul3(R, CA, tmp);
      // rcr = CB + tmp * R^T  (symmetric; build the upper part and mirror so the inverse is exactly symmetric)
      double u[6];
      u[0] = tmp[0] * R[0] + tmp[1] * R[1] + tmp[2] * R[2];
      u[1] = tmp[0] * R[3] + tmp[1] * R[4] + tmp[2] * R[5];
      u[2] = tmp[0] * R[6] + tmp[1] * R[7] + tmp[2] * R[8];
      u[3] = tmp[3] * R[3] + tmp[4] * R[4] + tmp[5] * R[5];
      u[4] = tmp[3] * R[6] + tmp[4] * R[7] + tmp[5] * R[8];
      u[5] = tmp[6] * R[6] + tmp[7] * R[7] + tmp[8] * R[8];
      rcr[0] = cb[0] + u[0]; rcr[1] = cb[1] + u[1]; rcr[2] = cb[2] + u[2];
      rcr[4] = cb[3] + u[3]; rcr[5] = cb[4] + u[4]; rcr[8] = cb[5] + u[5];
      rcr[3] = rcr[1]; rcr[6] = rcr[2]; rcr[7] = rcr[5];
      inv3(rcr, M);
      double* mo = A.mahal + (size_t)s * 6;
      mo[0] = M[0]; mo[1] = M[1]; mo[2] = M[2]; mo[3] = M[4]; mo[4] = M[5]; mo[5] = M[8];
      const float4 tb = A.tgt.sp[best_pos];
      const double ax = (double)p.x, ay = (double)p.y, az = (double)p.z;
      const double tx = P.T[0] * ax + P.T[1] * ay + P.T[2] * az + P.T[3];
      const double ty = P.T[4] * ax + P.T[5] * ay + P.T[6] * az + P.T[7];
      const double tz = P.T[8] * ax + P.T[9] * ay + P.T[10] * az + P.T[11];
      const double ex = (double)tb.x - tx, ey = (double)tb.y - ty, ez = (double)tb.z - tz;
      const double m00 = M[0], m01 = M[1], m02 = M[2], m11 = M[4], m12 = M[5], m22 = M[8];
      const double Mex = m00 * ex + m01 * ey + m02 * ez;
      const double Mey = m01 * ex + m11 * ey + m12 * ez;
      const double Mez = m02 * ex + m12 * ey + m22 * ez;
      acc[27] = ex * Mex + ey * Mey + ez * Mez;
      // S = skew(tA) = [[0,-tz,ty],[tz,0,-tx],[-ty,tx,0]];  MS = M*S
      const double ms00 = m01 * tz - m02 * ty, ms01 = -m00 * tz + m02 * tx, ms02 = m00 * ty - m01 * tx;
      const double ms10 = m11 * tz - m12 * ty, ms11 = -m01 * tz + m12 * tx, ms12 = m01 * ty - m11 * tx;
      const double ms20 = m12 * tz - m22 * ty, ms21 = -m02 * tz + m22 * tx, ms22 = m02 * ty - m12 * tx;
      // S^T * MS (upper): S^T = [[0,tz,-ty],[-tz,0,tx],[ty,-tx,0]]
      acc[0] = tz * ms10 - ty * ms20;   // (0,0)
      acc[1] = tz * ms11 - ty * ms21;   // (0,1)
      acc[2] = tz * ms12 - ty * ms22;   // (0,2)
      acc[3] = -ms00;                   // (0,3) = -(MS)[0][0]
      acc[4] = -ms10;                   // (0,4) = -(MS)[1][0]
      acc[5] = -ms20;                   // (0,5)
      acc[6] = -tz * ms01 + tx * ms21;  // (1,1)
      acc[7] = -tz * ms02 + tx * ms22;  // (1,2)
      acc[8] = -ms01;                   // (1,3)
      acc[9] = -ms11;                   // (1,4)
      acc[10] = -ms21;                  // (1,5)
      acc[11] = ty * ms02 - tx * ms12;  // (2,2)
      acc[12] = -ms02;                  // (2,3)
      acc[13] = -ms12;                  // (2,4)
      acc[14] = -ms22;                  // (2,5)
      acc[15] = m00; acc[16] = m01; acc[17] = m02;  // (3,3..5)
      acc[18] = m11; acc[19] = m12;                 // (4,4..5)
      acc[20] = m22;                                // (5,5)
      // b = J^T M e = [S^T Me ; -Me]
      acc[21] = tz * Mey - ty * Mez;
      acc[22] = -tz * Mex + tx * Mez;
      acc[23] = ty * Mex - tx * Mey;
      acc[24] = -Mex; acc[25] = -Mey; acc[26] = -Mez;
    }
  }
  block_reduce<kAcc>(acc, red);
  finish_partials<kAcc>(acc, A.partials, A.out, A.counter, A.flag, A.seq);
}

struct ErrArgs {
  const float4* ssp;
  int n_sorted;      // nleaf * 32 of the source
  const float4* tsp;
  const int* cpos;
  const double* mahal;
  double* partials;  // [blocks]
  double* out;       // [1]
  unsigned int* counter;
  unsigned long long* flag;
  unsigned long long seq;
};

constexpr int kErrThreads = 256;
__global__ void __launch_bounds__(kErrThreads) k_gicp_error(ErrArgs A, PoseArg P) {
  __shared__ double red[32];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[1] = {0.0};
  if (s < A.n_sorted) {
    const float4 p = A.ssp[s];
    int tp = (idx_bits(p.w) != kPadIdx) ? A.cpos[s] : -1;
    if (tp >= 0) {
      const float4 tb = A.tsp[tp];
      const double* m = A.mahal + (size_t)s * 6;
      const double ax = (double)p.x, ay = (double)p.y, az = (double)p.z;
      const double ex = (double)tb.x - (P.T[0] * ax + P.T[1] * ay + P.T[2] * az + P.T[3]);
      const double ey = (double)tb.y - (P.T[4] * ax + P.T[5] * ay + P.T[6] * az + P.T[7]);
      const double ez = (double)tb.z - (P.T[8] * ax + P.T[9] * ay + P.T[10] * az + P.T[11]);
      const double Mex = m[0] * ex + m[1] * ey + m[2] * ez;
      const double Mey = m[1] * ex + m[3] * ey + m[4] * ez;
      const double Mez = m[2] * ex + m[4] * ey + m[5] * ez;
      acc[0] = ex * Mex + ey * Mey + ez * Mez;
    }
  }
  block_reduce<1>(acc, red);
  finish_partials<1>(acc, A.partials, A.out, A.counter, A.flag, A.seq);
}

}  // namespace b2r
