// tests/lm_harness.cu — CPU-side check of the device-resident Levenberg-Marquardt state machine (csrc/pair_engine.cuh).
//
// The engine runs fast_gicp's step_lm loop inside k_pair_accumulate's last block: lm_advance() consumes the 29 reduced values of
// a round and decides what the next round evaluates.  Here the SAME functions (they are __host__ __device__) are driven on the
// CPU, with the oracle's orc_gicp_linearize / orc_gicp_error standing in for the two kernels of a round, and the outcome is
// compared with the oracle's own orc_gicp_align: iteration count, convergence flag, final pose, last correspondences.
// Test infrastructure: built by tests/test_host_logic.py with nvcc (host code only), never shipped.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include "../hdl_graph_slam_b200/csrc/engine.cuh"
#include "../hdl_graph_slam_b200/csrc/pair_engine.cuh"
#include "../oracle/oracle.h"
namespace b2r { thread_local std::string g_last_error; }
using namespace b2r;

static std::vector<float> read_f32(const char* path) {
  std::vector<float> v;
  FILE* f = fopen(path, "rb");
  if (!f) return v;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  v.resize(sz / 4);
  if (fread(v.data(), 4, v.size(), f) != v.size()) v.clear();
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  // lm_harness src.f32 tgt.f32 stride max_iterations trans_eps max_corr_dist g0..g15 (row-major float guess)
  if (argc < 7 + 16) { fprintf(stderr, "usage\n"); return 2; }
  std::vector<float> src = read_f32(argv[1]), tgt = read_f32(argv[2]);
  const size_t stride = (size_t)atoi(argv[3]);
  const size_t n = src.size() / stride, m = tgt.size() / stride;
  orc_gicp_config oc;
  oc.max_iterations = atoi(argv[4]);
  oc.transformation_epsilon = atof(argv[5]);
  oc.rotation_epsilon = 2e-3;
  oc.max_corr_dist = atof(argv[6]);
  oc.k_correspondences = 20;
  oc.num_threads = 0;
  float guess[16];
  for (int i = 0; i < 16; i++) guess[i] = (float)atof(argv[7 + i]);
  std::vector<double> scov(n * 9), tcov(m * 9);
  orc_gicp_covariances(src.data(), n, stride, 20, scov.data(), 0);
  orc_gicp_covariances(tgt.data(), m, stride, 20, tcov.data(), 0);
  // the oracle's own align
  orc_gicp_result ores;
  std::vector<int32_t> ocorr(n);
  orc_gicp_align(src.data(), n, stride, scov.data(), tgt.data(), m, stride, tcov.data(), &oc, guess, &ores, nullptr, nullptr, nullptr, ocorr.data());

  // the engine's state machine, kernels emulated by the oracle's linearize / error
  LmCfg cfg;
  cfg.max_iterations = oc.max_iterations; cfg.rot_eps = oc.rotation_epsilon; cfg.trans_eps = oc.transformation_epsilon;
  cfg.thr2 = oc.max_corr_dist * oc.max_corr_dist; cfg.lim = (float)cfg.thr2; cfg.want_fitness = 0; cfg.fit_max_range = 0; cfg.fit_lim = 0; cfg.index_seed = 0;
  PairDev P;
  std::memset(&P, 0, sizeof(P));
  PairReport rep;
  std::memset(&rep, 0, sizeof(rep));
  P.report = &rep;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 4; c++) { P.xe[r * 4 + c] = (double)guess[r * 4 + c]; P.x0[r * 4 + c] = (double)guess[r * 4 + c]; }
  P.mode = PM_FIRST; P.lambda = -1.0; P.nu = 2.0;
  std::vector<int32_t> corr[2] = {std::vector<int32_t>(n), std::vector<int32_t>(n)};
  std::vector<double> mahal[2] = {std::vector<double>(n * 9), std::vector<double>(n * 9)};
  std::vector<float> d2(n);
  int rounds = 0;
  while (P.mode != PM_DONE && rounds < 2000) {
    double T[16];
    for (int i = 0; i < 12; i++) T[i] = P.xe[i];
    T[12] = T[13] = T[14] = 0; T[15] = 1;
    double r[kAcc];
    for (int i = 0; i < kAcc; i++) r[i] = 0;
    const int mode = P.mode, cur = P.cur, wset = (mode == PM_FIRST) ? cur : (cur ^ 1);
    if (mode == PM_FUSED || mode == PM_ERR)
      r[28] = orc_gicp_error(src.data(), n, stride, tgt.data(), stride, corr[cur].data(), mahal[cur].data(), T, 0);
    if (mode == PM_FIRST || mode == PM_FUSED) {
      double H[36], b[6];
      r[27] = orc_gicp_linearize(src.data(), n, stride, scov.data(), tgt.data(), m, stride, tcov.data(), T, oc.max_corr_dist, corr[wset].data(), d2.data(),
                                 mahal[wset].data(), H, b, 0);
      int k = 0;
      for (int rr = 0; rr < 6; rr++)
        for (int cc = rr; cc < 6; cc++) r[k++] = H[rr * 6 + cc];
      for (int i = 0; i < 6; i++) r[21 + i] = b[i];
    }
    lm_advance(P, r, cfg);
    rounds++;
  }
  int bad = 0;
  if (P.mode != PM_DONE) { printf("state machine did not terminate\n"); bad++; }
  if (rep.r.iterations != ores.iterations) { printf("iterations %d vs oracle %d\n", rep.r.iterations, ores.iterations); bad++; }
  if (rep.r.converged != ores.converged) { printf("converged %d vs oracle %d\n", rep.r.converged, ores.converged); bad++; }
  if (rep.lm_failed != ores.lm_failed) { printf("lm_failed %d vs oracle %d\n", rep.lm_failed, ores.lm_failed); bad++; }
  double maxd = 0;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) maxd = std::fmax(maxd, std::fabs((double)rep.r.T[c * 4 + r] - (double)ores.T[r * 4 + c]));
  if (!(maxd <= 1e-6)) { printf("pose differs by %g\n", maxd); bad++; }
  size_t cdiff = 0;
  for (size_t i = 0; i < n; i++) cdiff += corr[rep.cur][i] != ocorr[i];
  if (cdiff) { printf("last correspondences differ in %zu points\n", cdiff); bad++; }
  // total LM trials: every round after the first is one trial
  if (rounds - 1 != ores.total_inner) { printf("trials %d vs oracle %d\n", rounds - 1, ores.total_inner); bad++; }
  printf("rounds=%d iterations=%d converged=%d max_pose_diff=%.3g corr_diff=%zu mismatches=%d\n", rounds, rep.r.iterations, rep.r.converged, maxd, cdiff, bad);
  return bad ? 1 : 0;
}
