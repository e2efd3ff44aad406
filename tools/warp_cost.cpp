// tools/warp_cost.cpp — offline work model of the search kernels on the CPU warp emulator (tests/warp_emu.hpp): runs the engine's
// traversal source for EVERY warp of a real synthetic scan pair and reports the per-warp work counters (leaf tests, tile visits,
// cooperative steps, collectives) whose maximum sets the kernel time on the GPU (profiles/r01_g_per_warp_profiles.md).
// Development aid: lets traversal heuristics be compared for their tail without a GPU.  Built and driven by tools/warp_cost.py.
//
//   warp_cost <target.f32> <source.f32> <n> <pose12.txt> <copies> <mode>     mode: 0 = 1-NN unseeded, 1 = 1-NN seeded by the answer's
//                                                                                  neighbour + hint (steady-state iteration), 2 = 20-NN self
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>
#include "../tests/warp_emu.hpp"
#define B2R_WARP_EMU 1
#define B2R_KNN_PROFILE 1
static long g_visit_hist[33];
#include "../hdl_graph_slam_b200/csrc/common.cuh"
#include "../hdl_graph_slam_b200/csrc/bvh.cuh"

using namespace b2r;

#include "../tests/host_bvh.hpp"

struct EmuKnn {
  static constexpr int kTileLanes = 3;
  static constexpr int kTileUnroll = 1;
  static constexpr bool kTwoPhase = true;
  static constexpr int K = 20;
  unsigned long long key[K];
  int n_tile = 0, n_coop = 0, n_try = 0, n_ins = 0;
  void reset() { for (int j = 0; j < K; j++) key[j] = kKeyInf; }
  float worst() const { return nn_key_d2(key[K - 1]); }
  float limit() const { return INFINITY; }
  void visit(float d2, int idx, int) {
    const unsigned long long kq = nn_key(d2, idx);
    if (kq >= key[K - 1]) return;
    n_ins++;
    bool pj = true;
    for (int j = K - 1; j > 0; j--) { const unsigned long long lo = key[j - 1]; const bool pl = lo > kq; if (pj) key[j] = pl ? lo : kq; pj = pl; }
    if (pj) key[0] = kq;
  }
};

static std::vector<float> load(const char* path, int n) {
  std::vector<float> v((size_t)n * 4);
  FILE* f = fopen(path, "rb");
  if (!f || fread(v.data(), sizeof(float), v.size(), f) != v.size()) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return v;
}

struct Row { int warp; long coll; int tile, coop, tries, ins; double cost; };

static void report(std::vector<Row>& rows, const char* what) {
  auto pct = [&](std::vector<double> v, double p) { std::sort(v.begin(), v.end()); return v[(size_t)std::min<double>(v.size() - 1, p * v.size())]; };
  std::vector<double> c, t, co, tr, cl;
  for (auto& r : rows) { c.push_back(r.cost); t.push_back(r.tile); co.push_back(r.coop); tr.push_back(r.tries); cl.push_back((double)r.coll); }
  auto line = [&](const char* n, std::vector<double>& v) {
    double m = 0; for (double x : v) m += x; m /= v.size();
    printf("  %-12s mean %9.1f  p50 %8.0f  p90 %8.0f  p99 %8.0f  p99.9 %8.0f  max %8.0f\n", n, m, pct(v, 0.5), pct(v, 0.9), pct(v, 0.99), pct(v, 0.999), pct(v, 1.0));
  };
  printf("%s: %zu warps\n", what, rows.size());
  line("model_cycles", c); line("tile_visits", t); line("coop_steps", co); line("leaf_tests", tr); line("collectives", cl);
  std::sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.cost > b.cost; });
  printf("  heaviest:");
  for (int i = 0; i < 6 && i < (int)rows.size(); i++) printf(" [w%d cost %.0f tile %d coop %d try %d coll %ld]", rows[i].warp, rows[i].cost, rows[i].tile, rows[i].coop, rows[i].tries, rows[i].coll);
  printf("\n");
}

template <int C>
static void run_1nn(const HostBvh& T, const HostBvh& S, const float* Tf, int mode, int stride_warps) {
  constexpr int Q = 32 / C;
  const int nwarps = (S.b.nleaf * kLeaf) / Q;
  std::vector<Row> rows;
  struct Ex { float nnmax, gdiag, range; };
  std::vector<Ex> extra;
  static float nn_d[32], gq[32][3], g_nnmax, g_gdiag, g_range;
  static bool gact[32];
  // steady-state seeds: the exact answer of a slightly different pose (like the previous LM iteration)
  float Tp[12];
  for (int i = 0; i < 12; i++) Tp[i] = Tf[i];
  Tp[3] += 0.03f; Tp[7] -= 0.02f;
  for (int w = 0; w < nwarps; w += stride_warps) {
    int seed_pos[32];
    if (mode == 1) {
      for (int l = 0; l < 32; l++) {
        const float4 p = S.sp[w * Q + (l & (Q - 1))];
        seed_pos[l] = -1;
        if (idx_bits(p.w) == kPadIdx) continue;
        Nn1 v; v.reset(6.25f);
        bvh_search_one(T.b, xform_row(Tp[0], Tp[1], Tp[2], Tp[3], p.x, p.y, p.z), xform_row(Tp[4], Tp[5], Tp[6], Tp[7], p.x, p.y, p.z),
                       xform_row(Tp[8], Tp[9], Tp[10], Tp[11], p.x, p.y, p.z), v);
        if (v.best_pos >= 0 && v.best_d2() < 6.25f) seed_pos[l] = v.best_pos;
      }
    }
    int tile = 0, coop = 0, tries = 0;
    wemu::Warp* wp = nullptr;
    long coll = 0;
    wemu::run_warp(0, [&](int l) {
      const float4 p = S.sp[w * Q + (l & (Q - 1))];
      const bool act = idx_bits(p.w) != kPadIdx;
      const float qx = xform_row(Tf[0], Tf[1], Tf[2], Tf[3], p.x, p.y, p.z), qy = xform_row(Tf[4], Tf[5], Tf[6], Tf[7], p.x, p.y, p.z),
                  qz = xform_row(Tf[8], Tf[9], Tf[10], Tf[11], p.x, p.y, p.z);
      Nn1 v;
      v.reset(6.25f);
      int sp0 = -1;
      if (mode == 1 && act && seed_pos[l] >= 0) {
        sp0 = seed_pos[l];
        const float4 t = T.sp[sp0];
        v.seed(dist2_f32(qx, qy, qz, t.x, t.y, t.z), idx_bits(t.w), sp0);
      }
      int hint = -1;
      const unsigned hm = __ballot_sync(0xffffffffu, sp0 >= 0);
      if (hm) hint = __shfl_sync(0xffffffffu, sp0, (int)__fns(hm, 0, (__popc(hm) + 1) / 2)) >> 5;
      bvh_group_search<C>(T.b, qx, qy, qz, act, v, -1, hint);
      if (l == 0) { tile = v.n_tile; coop = v.n_coop; tries = v.n_try; coll = (long)wemu::g()->ncoll[0]; }
      nn_d[l] = act ? std::sqrt(v.bd2) : 0.f;
      gq[l][0] = qx; gq[l][1] = qy; gq[l][2] = qz; gact[l] = act;
    });
    (void)wp;
    {
      float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
      g_nnmax = 0.f;
      for (int l = 0; l < Q; l++) if (gact[l]) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], gq[l][a]); hi[a] = std::max(hi[a], gq[l][a]); } if (std::isfinite(nn_d[l])) g_nnmax = std::max(g_nnmax, nn_d[l]); }
      g_gdiag = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
      g_range = std::sqrt(gq[0][0] * gq[0][0] + gq[0][1] * gq[0][1]);
      extra.push_back({g_nnmax, g_gdiag, g_range});
    }
    // cycles ~ fit of profiles/r01_g (seeded, 2 lanes per query), tile cost scaled by the lanes per query
    const double cost = 6400.0 + 3200.0 * (2.0 / C + 0.35) / 1.35 * tile + 440.0 * coop + 380.0 * tries;
    rows.push_back({w, coll, tile, coop, tries, 0, cost});
  }
  if (const char* dump = getenv("WARP_COST_DUMP")) {  // per-warp model cost + extent of the warp's source leaf, for scheduling studies
    if (FILE* f = fopen(dump, "w")) {
      for (auto& r : rows) {
        const int leaf = (r.warp * Q) / kLeaf;
        const float4 lo = S.llo[leaf], hi = S.lhi[leaf];
        const double diag = std::sqrt((double)(hi.x - lo.x) * (hi.x - lo.x) + (double)(hi.y - lo.y) * (hi.y - lo.y) + (double)(hi.z - lo.z) * (hi.z - lo.z));
        const Ex& e = extra[&r - &rows[0]];
        fprintf(f, "%d %.1f %d %.4f %.3f %.3f %.2f %d %d\n", r.warp, r.cost, leaf, diag, e.nnmax, e.gdiag, e.range, r.tile, r.tries);
      }
      fclose(f);
    }
  }
  report(rows, mode ? "1-NN seeded" : "1-NN unseeded");
  printf("  visited leaves by number of interested queries:");
  for (int i = 1; i <= Q; i++) printf(" %d:%ld", i, g_visit_hist[i]);
  printf("\n");
}

static void run_knn(const HostBvh& H, int stride_warps) {
  std::vector<Row> rows;
  for (int leaf = 0; leaf < H.b.nleaf; leaf += stride_warps) {
    int tile = 0, coop = 0, tries = 0, ins = 0;
    long coll = 0;
    int ins_l[32];
    wemu::run_warp(0, [&](int l) {
      const float4 q = H.sp[leaf * kLeaf + l];
      EmuKnn L;
      L.reset();
      bvh_group_search(H.b, q.x, q.y, q.z, idx_bits(q.w) != kPadIdx, L, leaf);
      ins_l[l] = L.n_ins;
      if (l == 0) { tile = L.n_tile; coop = L.n_coop; tries = L.n_try; coll = (long)wemu::g()->ncoll[0]; }
    });
    for (int l = 0; l < 32; l++) ins = std::max(ins, ins_l[l]);
    const double cost = 53500.0 + 5600.0 * tile + 3500.0 * coop + 730.0 * ins;  // fit of profiles/r01_g §1 (two-phase build)
    rows.push_back({leaf, coll, tile, coop, tries, ins, cost});
  }
  report(rows, "20-NN self");
}

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: warp_cost target.f32 source.f32 n pose12.txt copies mode [stride_warps]\n"); return 2; }
  const int n = atoi(argv[3]), copies = atoi(argv[5]), mode = atoi(argv[6]);
  const int stride = argc > 7 ? atoi(argv[7]) : 1;
  std::vector<float> tp = load(argv[1], n), sp = load(argv[2], n);
  float Tf[12];
  FILE* f = fopen(argv[4], "r");
  for (int i = 0; i < 12; i++) if (!f || fscanf(f, "%f", &Tf[i]) != 1) { fprintf(stderr, "bad pose file\n"); return 2; }
  fclose(f);
  HostBvh T = build(tp, n), S = build(sp, n);
  if (mode == 2) { run_knn(S, stride); return 0; }
  switch (copies) {
    case 1: run_1nn<1>(T, S, Tf, mode, stride); break;
    case 2: run_1nn<2>(T, S, Tf, mode, stride); break;
    case 4: run_1nn<4>(T, S, Tf, mode, stride); break;
    case 8: run_1nn<8>(T, S, Tf, mode, stride); break;
    default: return 2;
  }
  return 0;
}
