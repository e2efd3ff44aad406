"""CPU-only: the gating half of hdl_graph_slam's LoopDetector (include/hdl_graph_slam/loop_detector.hpp:39-46,57-68,81-109,137-142) behind the
C ABI — find_candidates, the initial guess, and detect() split into plan / batched matching / sequential replay — against an independent numpy /
scipy restatement of the reference's sequential code.  No GPU: the matching results are faked."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import hdl_graph_slam_b200 as pkg


def random_pose(rng, scale=30.0, noise=0.0):
    T = np.eye(4)
    T[:3, :3] = Rotation.from_rotvec(rng.normal(size=3) * 0.7).as_matrix()
    if noise:
        T[:3, :3] += rng.normal(size=(3, 3)) * noise  # a drifted (not exactly orthonormal) graph estimate
    T[:3, 3] = rng.uniform(-scale, scale, 3)
    return T


def ref_find_candidates(p, keyframes, new_kf, last_edge):
    """loop_detector.hpp:81-109 restated"""
    if new_kf[0] - last_edge < p["min_edge_interval"]:
        return []
    out = []
    for i, (acc, est) in enumerate(keyframes):
        if new_kf[0] - acc < p["accum_distance_thresh"]:
            continue
        if np.linalg.norm(est[:2, 3] - new_kf[1][:2, 3]) > p["distance_thresh"]:
            continue
        out.append(i)
    return out


def ref_guess(new_est, cand_est):
    """:137-142 with scipy: rotations through a normalised quaternion, inverse * candidate, float32, z = 0"""
    def renorm(T):
        R = T.copy()
        R[:3, :3] = Rotation.from_quat(Rotation.from_matrix(T[:3, :3]).as_quat()).as_matrix() if abs(np.linalg.det(T[:3, :3]) - 1) < 1e-9 else eigen_renorm(T[:3, :3])
        return R
    a, b = renorm(new_est), renorm(cand_est)
    ainv = np.eye(4)
    ainv[:3, :3] = a[:3, :3].T
    ainv[:3, 3] = -a[:3, :3].T @ a[:3, 3]
    g = (ainv @ b).astype(np.float32)
    g[2, 3] = 0.0
    return g


def eigen_renorm(m):
    """Eigen 3.3 Quaterniond(m).normalized().toRotationMatrix() for a matrix that is not exactly a rotation"""
    t = np.trace(m)
    q = np.zeros(4)  # x y z w
    if t > 0:
        t = np.sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t
        q[0] = (m[2, 1] - m[1, 2]) * t; q[1] = (m[0, 2] - m[2, 0]) * t; q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]: i = 1
        if m[2, 2] > m[i, i]: i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0); q[i] = 0.5 * t; t = 0.5 / t
        q[3] = (m[k, j] - m[j, k]) * t; q[j] = (m[j, i] + m[i, j]) * t; q[k] = (m[k, i] + m[i, k]) * t
    q /= np.linalg.norm(q)
    return Rotation.from_quat(q).as_matrix()


def trajectory(rng, n, step=2.0):
    """keyframes along a closed circuit: (accum_distance, estimate)"""
    out, acc = [], 0.0
    for k in range(n):
        ang = 2 * np.pi * k / n
        T = np.eye(4)
        T[:3, :3] = Rotation.from_euler("z", ang + np.pi / 2).as_matrix()
        T[:3, 3] = [20 * np.cos(ang) + rng.normal() * 0.3, 20 * np.sin(ang) + rng.normal() * 0.3, rng.normal() * 0.1]
        out.append((acc, T))
        acc += step
    return out


def test_defaults_are_the_rosparam_defaults():
    g = pkg.LoopClosureGate()
    p = g.params
    assert (p.distance_thresh, p.accum_distance_thresh, p.min_edge_interval, p.fitness_score_thresh) == (5.0, 8.0, 5.0, 0.5)
    assert p.fitness_score_max_range == np.finfo(np.float64).max
    assert g.last_edge_accum_distance == 0.0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_find_candidates_matches_the_reference_logic(seed):
    rng = np.random.default_rng(seed)
    kfs = trajectory(rng, 70)
    g = pkg.LoopClosureGate()
    p = dict(distance_thresh=5.0, accum_distance_thresh=8.0, min_edge_interval=5.0)
    hits = 0
    for last_edge in (0.0, 60.0, 139.0):
        g.last_edge_accum_distance = last_edge
        for k in range(5, 70, 3):
            new_kf = (kfs[k][0] + 140.0, kfs[k][1].copy())  # second lap: same places, 140 m later
            new_kf[1][:3, 3] += rng.normal(size=3) * 0.5
            want = ref_find_candidates(p, kfs, new_kf, last_edge)
            assert g.find_candidates(kfs, new_kf) == want
            hits += len(want)
    assert hits > 50
    # exact thresholds: '<' on the accumulated distances, '>' on the xy distance
    flat = [(0.0, np.eye(4))]
    g.last_edge_accum_distance = 0.0
    nk = np.eye(4); nk[0, 3] = 5.0
    assert g.find_candidates(flat, (8.0, nk)) == [0]      # 8 - 0 == accum thresh (not '<'), dist == 5 (not '>')
    assert g.find_candidates(flat, (7.999, nk)) == []
    nk2 = np.eye(4); nk2[0, 3] = 5.0001
    assert g.find_candidates(flat, (8.0, nk2)) == []
    g.last_edge_accum_distance = 3.0001
    assert g.find_candidates(flat, (8.0, nk)) == []       # 8 - 3.0001 < 5: too close to the last loop edge


@pytest.mark.parametrize("noise", [0.0, 1e-3])
def test_guess_matches_eigen_semantics(noise):
    rng = np.random.default_rng(5)
    g = pkg.LoopClosureGate()
    for _ in range(50):
        a, b = random_pose(rng, noise=noise), random_pose(rng, noise=noise)
        got = g.guess(a, b)
        want = ref_guess(a, b)
        assert got.dtype == np.float32 and got[2, 3] == 0.0 and np.array_equal(got[3], [0, 0, 0, 1])
        assert np.max(np.abs(got - want)) < 2e-6 * max(1.0, np.max(np.abs(want)))
    # a rotation of ~180 degrees about each axis takes the non-positive-trace branches of the quaternion conversion
    for axis in "xyz":
        a = np.eye(4); a[:3, :3] = Rotation.from_euler(axis, 179.5, degrees=True).as_matrix(); a[:3, 3] = [1, 2, 3]
        assert np.max(np.abs(g.guess(a, np.eye(4)) - ref_guess(a, np.eye(4)))) < 2e-6 * 4


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_plan_and_replay_equal_the_sequential_detect(seed):
    """detect() of the reference, run sequentially with a fake matching (a deterministic function of the candidate set), must equal
    plan -> (fake batch) -> replay, including the updates of last_edge_accum_distance between new keyframes"""
    rng = np.random.default_rng(seed)
    kfs = trajectory(rng, 60)
    new = []
    for k in sorted(rng.choice(60, 14, replace=False)):
        T = kfs[k][1].copy(); T[:3, 3] += rng.normal(size=3) * 0.4
        new.append((kfs[k][0] + 120.0 + rng.uniform(0, 0.5), T))
    p = dict(distance_thresh=5.0, accum_distance_thresh=8.0, min_edge_interval=5.0)

    def fake_matching(g_index, cands):  # index inside the group or -1, as b2r_batch_loop_detect reports it
        if not cands:
            return -1
        h = (g_index * 7 + sum(cands)) % 5
        return -1 if h == 0 else h % len(cands)

    # the reference's sequential walk
    last = float(rng.uniform(0, 100))
    start_last = last
    want = []
    for gi, nk in enumerate(new):
        cands = ref_find_candidates(p, kfs, nk, last)
        b = fake_matching(gi, cands)
        if b >= 0:
            want.append((gi, cands[b]))
            last = nk[0]
    # plan / batch / replay
    gate = pkg.LoopClosureGate()
    gate.last_edge_accum_distance = start_last
    cand, guesses, gf = gate.plan(kfs, new)
    assert len(gf) == len(new) + 1 and gf[0] == 0 and gf[-1] == len(cand) == len(guesses)
    best = []
    for gi in range(len(new)):
        group = cand[gf[gi]:gf[gi + 1]]
        assert group == ref_find_candidates(p, kfs, new[gi], start_last)
        for j, c in enumerate(group):
            assert np.array_equal(guesses[gf[gi] + j], gate.guess(new[gi][1], kfs[c][1]))
        best.append(fake_matching(gi, group))
    accepted = gate.replay(new, gf, best, start_last)
    got = [(gi, cand[gf[gi] + a]) for gi, a in enumerate(accepted) if a >= 0]
    assert got == want
    assert gate.last_edge_accum_distance == last
    assert len(want) >= 1


def test_argument_errors_are_reported_not_crashed():
    import ctypes as C
    from hdl_graph_slam_b200 import _capi
    lib = _capi.load()
    assert lib.b2r_loop_params_default(None) == _capi.B2R_EINVAL
    p = _capi.LoopParams(); lib.b2r_loop_params_default(C.byref(p))
    ks = (_capi.KeyframeState * 2)()
    for i in range(2):
        ks[i].accum_distance = 0.0
        for k in (0, 5, 10, 15):
            ks[i].estimate[k] = 1.0
    nk = _capi.KeyframeState(); nk.accum_distance = 100.0
    for k in (0, 5, 10, 15):
        nk.estimate[k] = 1.0
    n = C.c_size_t()
    out = (C.c_int32 * 1)()
    assert lib.b2r_loop_find_candidates(C.byref(p), ks, 2, C.byref(nk), 0.0, out, 1, C.byref(n)) == _capi.B2R_EINVAL and n.value == 2  # buffer too small
    assert b"too small" in lib.b2r_last_error()
