// loop_gate.cuh — the host-side gating of hdl_graph_slam's LoopDetector around the batched registration (no device code).
//
// Mirrors /root/reference/include/hdl_graph_slam/loop_detector.hpp:
//   b2r_loop_params_default   <- constructor's rosparams                                     :39-46
//   b2r_loop_find_candidates  <- LoopDetector::find_candidates                               :81-109
//   b2r_loop_guess            <- the initial guess of LoopDetector::matching                 :137-142
//   b2r_loop_detect_plan / b2r_loop_detect_replay <- LoopDetector::detect                    :57-68
// detect() walks the new keyframes one by one: find_candidates (whose first gate looks at the accumulated distance of the LAST
// REGISTERED loop edge), matching, and — if a loop is found — the update of that distance (:166).  The only coupling between
// two new keyframes is that scalar, and matching() has no other side effect, so the walk is split in three:
//   plan    every new keyframe's candidates under the gate evaluated with the distance the walk STARTS with (the weakest
//           the gate can be during the walk whenever the accumulated distances do not decrease along the walk; replay is
//           exact in any case, see below) — all (candidate, new keyframe, guess) pairs of the walk, grouped per new keyframe;
//   batch   b2r_batch_loop_detect aligns every pair (sharded over the GPUs) and returns each group's best candidate;
//   replay  the sequential walk itself on those results: a new keyframe whose gate fails with the distance as it stands at
//           ITS turn reports no loop (its speculative matching is dropped), an accepted loop moves the distance.
// A keyframe that passes the gate at its turn but was planned without candidates because the STARTING distance gated it out
// cannot occur when the distance only grows; replay reports it (return value B2R_ESTATE) instead of guessing.
#pragma once
#include <cmath>
#include <cfloat>
#include <cstring>
#include "engine.cuh"

namespace b2r {

// Eigen::Quaterniond(R).normalized().toRotationMatrix() on the rotation block of a column-major 4x4 (Eigen 3.3's
// quaternionbase_assign_impl<..., 3, 3> and QuaternionBase::toRotationMatrix, restated)
inline void renormalized_rotation(const double* M /* column-major 4x4 */, double R[9] /* row-major */) {
  auto m = [&](int r, int c) { return M[c * 4 + r]; };
  double q[4];  // x, y, z, w
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

inline bool loop_first_gate(const b2r_loop_params& p, double new_accum_distance, double last_edge_accum_distance) {
  return !(new_accum_distance - last_edge_accum_distance < p.min_edge_interval);  // :83 "too close to the last registered loop edge"
}

// candidates of one new keyframe past the first gate (:89-106), in the order of `keyframes`; returns the count, writes at most `capacity`
inline size_t loop_candidates_of(const b2r_loop_params& p, const b2r_keyframe_state* keyframes, size_t n, const b2r_keyframe_state& nk, int32_t* out, size_t capacity) {
  size_t cnt = 0;
  for (size_t i = 0; i < n; i++) {
    const b2r_keyframe_state& k = keyframes[i];
    if (nk.accum_distance - k.accum_distance < p.accum_distance_thresh) continue;  // :91 travelled distance between the keyframes too small
    const double dx = k.estimate[12] - nk.estimate[12], dy = k.estimate[13] - nk.estimate[13];
    const double dist = std::sqrt(dx * dx + dy * dy);  // :99 (pos1.head<2>() - pos2.head<2>()).norm()
    if (dist > p.distance_thresh) continue;            // :100
    if (cnt < capacity && out) out[cnt] = (int32_t)i;
    cnt++;
  }
  return cnt;
}

}  // namespace b2r

extern "C" int b2r_loop_params_default(b2r_loop_params* p) {
  if (!p) return b2r::fail(B2R_EINVAL, "NULL argument");
  p->distance_thresh = 5.0;          // :40
  p->accum_distance_thresh = 8.0;    // :41
  p->min_edge_interval = 5.0;        // :42
  p->fitness_score_max_range = DBL_MAX;  // :44 std::numeric_limits<double>::max()
  p->fitness_score_thresh = 0.5;     // :45
  return B2R_OK;
}

extern "C" int b2r_loop_find_candidates(const b2r_loop_params* p, const b2r_keyframe_state* keyframes, size_t n_keyframes, const b2r_keyframe_state* new_keyframe,
                                        double last_edge_accum_distance, int32_t* candidates, size_t capacity, size_t* n_candidates) {
  if (!p || !new_keyframe || !n_candidates || (n_keyframes && !keyframes)) return b2r::fail(B2R_EINVAL, "NULL argument");
  *n_candidates = 0;
  if (!b2r::loop_first_gate(*p, new_keyframe->accum_distance, last_edge_accum_distance)) return B2R_OK;
  *n_candidates = b2r::loop_candidates_of(*p, keyframes, n_keyframes, *new_keyframe, candidates, capacity);
  if (*n_candidates > capacity) return b2r::fail(B2R_EINVAL, "candidate buffer too small (n_candidates holds the number needed)");
  return B2R_OK;
}

extern "C" int b2r_loop_guess(const double* new_keyframe_estimate, const double* candidate_estimate, float* guess) {
  if (!new_keyframe_estimate || !candidate_estimate || !guess) return b2r::fail(B2R_EINVAL, "NULL argument");
  double Rn[9], Rc[9];
  b2r::renormalized_rotation(new_keyframe_estimate, Rn);  // :138
  b2r::renormalized_rotation(candidate_estimate, Rc);     // :140
  const double* tn = new_keyframe_estimate + 12;
  const double* tc = candidate_estimate + 12;
  // new^-1 * cand for isometries: [Rn^T Rc | Rn^T (tc - tn)]   (:141)
  double G[16] = {0};
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += Rn[k * 3 + r] * Rc[k * 3 + c];
      G[c * 4 + r] = s;
    }
    // Eigen composes inverse() first: translation of the inverse = -(Rn^T tn), then adds Rn^T tc
    double inv_t = 0.0, rt = 0.0;
    for (int k = 0; k < 3; k++) { inv_t += Rn[k * 3 + r] * tn[k]; rt += Rn[k * 3 + r] * tc[k]; }
    G[12 + r] = rt + (-inv_t);
  }
  G[15] = 1.0;
  for (int i = 0; i < 16; i++) guess[i] = (float)G[i];  // .matrix().cast<float>()
  guess[14] = 0.0f;                                      // :142 guess(2, 3) = 0.0
  return B2R_OK;
}

extern "C" int b2r_loop_detect_plan(const b2r_loop_params* p, const b2r_keyframe_state* keyframes, size_t n_keyframes, const b2r_keyframe_state* new_keyframes,
                                    size_t n_new, double last_edge_accum_distance, int32_t* candidate_index, float* guesses, size_t capacity,
                                    int64_t* group_first, size_t* n_pairs) {
  if (!p || !group_first || !n_pairs || (n_keyframes && !keyframes) || (n_new && !new_keyframes)) return b2r::fail(B2R_EINVAL, "NULL argument");
  size_t total = 0;
  group_first[0] = 0;
  for (size_t g = 0; g < n_new; g++) {
    const b2r_keyframe_state& nk = new_keyframes[g];
    if (b2r::loop_first_gate(*p, nk.accum_distance, last_edge_accum_distance)) {
      const size_t room = total < capacity ? capacity - total : 0;
      const size_t cnt = b2r::loop_candidates_of(*p, keyframes, n_keyframes, nk, candidate_index ? candidate_index + total : nullptr, candidate_index ? room : 0);
      if (guesses && candidate_index)
        for (size_t j = 0; j < cnt && total + j < capacity; j++) {
          const int rc = b2r_loop_guess(nk.estimate, keyframes[candidate_index[total + j]].estimate, guesses + (total + j) * 16);
          if (rc) return rc;
        }
      total += cnt;
    }
    group_first[g + 1] = (int64_t)total;
  }
  *n_pairs = total;
  if (total > capacity) return b2r::fail(B2R_EINVAL, "pair buffers too small (n_pairs holds the number needed)");
  return B2R_OK;
}

extern "C" int b2r_loop_detect_replay(const b2r_loop_params* p, const b2r_keyframe_state* new_keyframes, size_t n_new, const int64_t* group_first, const int32_t* best,
                                      double planned_last_edge_accum_distance, double* last_edge_accum_distance, int32_t* accepted) {
  if (!p || !group_first || !last_edge_accum_distance || (n_new && (!new_keyframes || !best || !accepted))) return b2r::fail(B2R_EINVAL, "NULL argument");
  double last = *last_edge_accum_distance;
  int rc = B2R_OK;
  for (size_t g = 0; g < n_new; g++) {
    accepted[g] = -1;
    const double acc = new_keyframes[g].accum_distance;
    if (!b2r::loop_first_gate(*p, acc, last)) continue;  // gated at ITS turn: the reference would not even have called matching()
    if (!b2r::loop_first_gate(*p, acc, planned_last_edge_accum_distance)) {
      // passes now but was planned out: only possible if the distance went DOWN during the walk, which detect() never does
      rc = b2r::fail(B2R_ESTATE, "a new keyframe passes the loop-edge gate at its turn but was gated out of the plan");
      continue;
    }
    if (best[g] >= 0 && best[g] < (int32_t)(group_first[g + 1] - group_first[g])) {
      accepted[g] = best[g];
      last = acc;  // :166 last_edge_accum_distance = new_keyframe->accum_distance
    }
  }
  *last_edge_accum_distance = last;
  return rc;
}
