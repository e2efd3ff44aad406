// tests/warp_harness.cpp — CPU-side execution of the engine's WARP-GROUP traversal (csrc/bvh.cuh: bvh_group_search with its group
// masks, per-lane refinement, centred window, tile / two-phase / cooperative leaf visits and C lanes per query) on an emulated
// 32-lane warp (tests/warp_emu.hpp), checked bit-for-bit against the oracle's exact k-NN.  Test infrastructure: built with g++ by
// tests/test_host_logic.py, never shipped.
//
//   warp_harness <n_points> <n_groups> <mode> <copies> [target.f32 source.f32 pose x12]     (mode 3: two scans from files)
// 1-NN part mirrors k_gicp_correspond: a warp carries 32/copies consecutive (Hilbert-sorted) points of a SOURCE cloud, transformed
// by a pose; variants: unseeded, seeded with a real candidate (+ hinted start leaf), range-limited (GICP's 2.5 m), groups that
// straddle far-apart regions, inactive lanes.  k-NN part mirrors k_knn_cov: a warp = one leaf of the cloud against itself.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>
#include "warp_emu.hpp"
#define B2R_WARP_EMU 1
#include "../hdl_graph_slam_b200/csrc/common.cuh"
#include "../hdl_graph_slam_b200/csrc/bvh.cuh"
#include "../oracle/oracle.h"

using namespace b2r;

#include "host_bvh.hpp"
// list visitor with the interface of the engine's KnnList / KnnRegs (sorted packed (d2, idx) keys, two-phase tile visits)
struct EmuKnn {
  static constexpr int kTileLanes = 3;
  static constexpr int kTileUnroll = 1;
  static constexpr bool kTwoPhase = true;
  static constexpr int K = 20;
  unsigned long long key[K];
  void reset() { for (int j = 0; j < K; j++) key[j] = kKeyInf; }
  float worst() const { return nn_key_d2(key[K - 1]); }
  float limit() const { return INFINITY; }
  void visit(float d2, int idx, int) {
    const unsigned long long kq = nn_key(d2, idx);
    if (kq >= key[K - 1]) return;
    bool pj = true;  // same slot update as KnnRegs::visit
    for (int j = K - 1; j > 0; j--) {
      const unsigned long long lo = key[j - 1];
      const bool pl = lo > kq;
      if (pj) key[j] = pl ? lo : kq;
      pj = pl;
    }
    if (pj) key[0] = kq;
  }
};

// V = Nn1 (float distance + index + position) or Nn1K (one packed key, what k_pair_search carries)
static bool nn_found(const Nn1& v) { return v.best_pos >= 0; }
static bool nn_found(const Nn1K& v) { return v.found(); }
static bool nn_pos_ok(const HostBvh& T, const Nn1& v, int want) { return idx_bits(T.sp[v.best_pos].w) == want; }
static bool nn_pos_ok(const HostBvh&, const Nn1K&, int) { return true; }
template <int C, class V>
static long run_1nn(const HostBvh& T, const HostBvh& S, int n_src, const float* Tf, int n_groups, float lim, int variant, const std::vector<float>& tgt_pts,
                    int n_tgt, long* checked) {
  constexpr int Q = 32 / C;
  long bad = 0;
  const int total_groups = (S.b.nleaf * kLeaf) / Q;
  for (int gi = 0; gi < n_groups; gi++) {
    // spread the sampled groups over the cloud; variant 3 builds "straddling" groups out of two far-apart halves
    const int g0 = (int)(((long long)gi * 7919) % total_groups);
    const int g1 = (int)(((long long)gi * 104729 + total_groups / 2) % total_groups);
    float qx[32], qy[32], qz[32];
    bool act[32];
    int spos[32];
    for (int l = 0; l < 32; l++) {
      const int q = l & (Q - 1);
      const int grp = (variant == 3 && q >= Q / 2 && Q > 1) ? g1 : g0;
      const int s = grp * Q + q;
      spos[l] = s;
      const float4 p = S.sp[s];
      act[l] = idx_bits(p.w) != kPadIdx && !(variant == 4 && (q % 3) == 1);  // variant 4: a third of the lanes inactive
      qx[l] = xform_row(Tf[0], Tf[1], Tf[2], Tf[3], p.x, p.y, p.z);
      qy[l] = xform_row(Tf[4], Tf[5], Tf[6], Tf[7], p.x, p.y, p.z);
      qz[l] = xform_row(Tf[8], Tf[9], Tf[10], Tf[11], p.x, p.y, p.z);
    }
    // exact answers from the oracle
    std::vector<float> qs(32 * 4);
    for (int l = 0; l < 32; l++) { qs[l * 4] = qx[l]; qs[l * 4 + 1] = qy[l]; qs[l * 4 + 2] = qz[l]; qs[l * 4 + 3] = 1.f; }
    int32_t oi[32]; float od[32];
    orc_knn(tgt_pts.data(), n_tgt, 4, qs.data(), 32, 4, 1, oi, od, 1);
    // seeds (variants 1, 2): a REAL target point near the query but usually not the answer: the sorted neighbour of the answer
    int seed_pos[32];
    for (int l = 0; l < 32; l++) {
      seed_pos[l] = -1;
      if ((variant == 1 || variant == 2) && act[l] && (l % 5) != 0) {
        // position of the oracle's answer in the target's sorted array, shifted by a few slots
        int pos = -1;
        for (int s = 0; s < T.b.nleaf * kLeaf; s++) if (idx_bits(T.sp[s].w) == oi[l]) { pos = s; break; }
        int sp = pos + ((l % 3) - 1) * 3;
        if (sp < 0 || sp >= T.b.nleaf * kLeaf || idx_bits(T.sp[sp].w) == kPadIdx) sp = pos;
        seed_pos[l] = sp;
      }
    }
    V res[32];
    wemu::run_warp((unsigned)(gi % 4) * 32, [&](int l) {
      V v;
      v.reset(lim);
      int sp0 = seed_pos[l];
      if (act[l] && sp0 >= 0) {
        const float4 t = T.sp[sp0];
        v.seed(dist2_f32(qx[l], qy[l], qz[l], t.x, t.y, t.z), idx_bits(t.w), sp0);
      } else sp0 = -1;
      int hint = -1;
      if (variant == 2) {  // as k_gicp_correspond: the middle seeded lane's leaf
        const unsigned hm = __ballot_sync(0xffffffffu, sp0 >= 0);
        if (hm) hint = __shfl_sync(0xffffffffu, sp0, (int)__fns(hm, 0, (__popc(hm) + 1) / 2)) >> 5;
      }
      bvh_group_search<C>(T.b, qx[l], qy[l], qz[l], act[l], v, -1, hint);
      res[l] = v;
    });
    for (int l = 0; l < 32; l++) {
      if (!act[l]) continue;
      (*checked)++;
      const bool want_valid = od[l] < lim;
      const bool got_valid = nn_found(res[l]) && res[l].best_d2() < lim;
      bool ok = want_valid == got_valid;
      if (ok && want_valid) ok = res[l].best_idx() == oi[l] && res[l].best_d2() == od[l] && nn_pos_ok(T, res[l], oi[l]);
      // copies of one query must agree with each other exactly
      if (ok && C > 1) { const V& a = res[l & (Q - 1)]; ok = a.best_idx() == res[l].best_idx() && a.best_d2() == res[l].best_d2(); }
      if (!ok) {
        if (bad < 5) printf("1nn mismatch C=%d variant=%d group=%d lane=%d: got (%g,%d) want (%g,%d)\n", C, variant, gi, l, res[l].best_d2(), res[l].best_idx(), od[l], oi[l]);
        bad++;
      }
    }
  }
  return bad;
}

static long run_knn(const HostBvh& H, const std::vector<float>& pts, int n, int n_groups, long* checked) {
  long bad = 0;
  constexpr int K = EmuKnn::K;
  for (int gi = 0; gi < n_groups; gi++) {
    const int leaf = (int)(((long long)gi * 7919) % H.b.nleaf);
    std::vector<float> qs(32 * 4);
    bool act[32];
    for (int l = 0; l < 32; l++) {
      const float4 p = H.sp[leaf * kLeaf + l];
      act[l] = idx_bits(p.w) != kPadIdx;
      qs[l * 4] = act[l] ? p.x : 0.f; qs[l * 4 + 1] = act[l] ? p.y : 0.f; qs[l * 4 + 2] = act[l] ? p.z : 0.f; qs[l * 4 + 3] = 1.f;
    }
    std::vector<int32_t> oi(32 * K); std::vector<float> od(32 * K);
    orc_knn(pts.data(), n, 4, qs.data(), 32, 4, K, oi.data(), od.data(), 1);
    EmuKnn res[32];
    wemu::run_warp(0, [&](int l) {
      const float4 q = H.sp[leaf * kLeaf + l];
      EmuKnn L;
      L.reset();
      bvh_group_search(H.b, q.x, q.y, q.z, act[l], L, leaf);
      res[l] = L;
    });
    for (int l = 0; l < 32; l++) {
      if (!act[l]) continue;
      (*checked)++;
      bool ok = true;
      for (int j = 0; j < K && j < n; j++) {
        const int idx = (int)(unsigned)(res[l].key[j] & 0xffffffffull);
        if (idx != oi[l * K + j] || nn_key_d2(res[l].key[j]) != od[l * K + j]) { ok = false; break; }
      }
      if (!ok) { if (bad < 5) printf("knn mismatch leaf=%d lane=%d\n", leaf, l); bad++; }
    }
  }
  return bad;
}

template <int C>
static long all_1nn(const HostBvh& T, const HostBvh& S, int n_src, const float* Tf, int groups, const std::vector<float>& tp, int nt, long* checked) {
  long bad = 0;
  for (int variant = 0; variant <= 4; variant++)
    for (float lim : {INFINITY, 6.25f}) {
      bad += run_1nn<C, Nn1>(T, S, n_src, Tf, groups, lim, variant, tp, nt, checked);
      bad += run_1nn<C, Nn1K>(T, S, n_src, Tf, groups, lim, variant, tp, nt, checked);
    }
  return bad;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 6000;
  const int groups = argc > 2 ? atoi(argv[2]) : 24;
  const int mode = argc > 3 ? atoi(argv[3]) : 0;
  const int copies = argc > 4 ? atoi(argv[4]) : 4;
  unsigned long long s = 4242 + mode;
  std::vector<float> tp, sp((size_t)n * 4);
  float Tf[12];
  if (mode == 3) {
    // two real synthetic LiDAR scans (float32 x,y,z,1 records) written by the test: argv[5] target, argv[6] source, argv[7..18] pose rows
    if (argc < 19) { fprintf(stderr, "mode 3 needs target.f32 source.f32 and 12 pose floats\n"); return 2; }
    tp.resize((size_t)n * 4);
    FILE* ft = fopen(argv[5], "rb"); FILE* fs = fopen(argv[6], "rb");
    if (!ft || !fs || fread(tp.data(), 4, tp.size(), ft) != tp.size() || fread(sp.data(), 4, sp.size(), fs) != sp.size()) { fprintf(stderr, "cannot read the scans\n"); return 2; }
    fclose(ft); fclose(fs);
    for (int i = 0; i < 12; i++) Tf[i] = (float)atof(argv[7 + i]);
  } else {
    tp = make_points(n, mode, s);
    // the source cloud: the same surface sampled again (jittered), so that most queries have a neighbour within range
    for (int i = 0; i < n; i++) {
      const double j = mode == 1 ? 0.0 : 0.15;
      sp[i * 4] = tp[i * 4] + (float)((urand(s) - 0.5) * j); sp[i * 4 + 1] = tp[i * 4 + 1] + (float)((urand(s) - 0.5) * j);
      sp[i * 4 + 2] = tp[i * 4 + 2] + (float)((urand(s) - 0.5) * j); sp[i * 4 + 3] = 1.f;
    }
    const float c = std::cos(0.02f), sn = std::sin(0.02f);
    const float T0[12] = {c, -sn, 0.f, 0.35f, sn, c, 0.f, -0.2f, 0.f, 0.f, 1.f, 0.05f};
    for (int i = 0; i < 12; i++) Tf[i] = T0[i];
  }
  HostBvh T = build(tp, n), S = build(sp, n);
  long checked = 0, bad1 = 0;
  switch (copies) {
    case 1: bad1 = all_1nn<1>(T, S, n, Tf, groups, tp, n, &checked); break;
    case 2: bad1 = all_1nn<2>(T, S, n, Tf, groups, tp, n, &checked); break;
    case 4: bad1 = all_1nn<4>(T, S, n, Tf, groups, tp, n, &checked); break;
    case 8: bad1 = all_1nn<8>(T, S, n, Tf, groups, tp, n, &checked); break;
    default: fprintf(stderr, "copies must be 1, 2, 4 or 8\n"); return 2;
  }
  long checkedk = 0;
  const long badk = run_knn(T, tp, n, groups, &checkedk);
  printf("copies=%d mode=%d  1nn_checked=%ld 1nn_mismatch=%ld  knn_checked=%ld knn_mismatch=%ld\n", copies, mode, checked, bad1, checkedk, badk);
  return (bad1 || badk) ? 1 : 0;
}
