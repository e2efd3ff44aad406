"""GPU: the batched, device-resident registration path (b2r_batch_*, csrc/pair_engine.cuh) — many (source, target) pairs per
launch with the LM step on the device — gives, pair by pair, BITWISE the results of running each pair alone through the
pcl::Registration-shaped handle, whatever the batch composition, chunking or lanes-per-query; and matches the oracle.
Reference site: the candidate loop of LoopDetector::matching, include/hdl_graph_slam/loop_detector.hpp:135-154."""
import os
import numpy as np
import pytest
import hdl_graph_slam_b200 as pkg
from common import rot_err, trans_err, perturb

pytestmark = pytest.mark.gpu


def make_groups(synth, sensor="vlp16_16k", targets=(0, 10, 20), cands=(1, 251, 2, 252)):
    clouds, pairs, group_first = {}, [], [0]
    for gi, tf in enumerate(targets):
        clouds.setdefault(tf, synth.scan(sensor, frame=tf, stride=8))
        for j, off in enumerate(cands):
            sf = tf + off
            clouds.setdefault(sf, synth.scan(sensor, frame=sf, stride=8))
            rel = np.linalg.inv(synth.pose_matrix(tf)) @ synth.pose_matrix(sf)
            g = (rel @ perturb(60 + 7 * gi + j, 0.3, 2.0)).astype(np.float32)
            g[2, 3] = 0.0  # loop_detector.hpp:142
            pairs.append((sf, tf, g))
        group_first.append(len(pairs))
    return clouds, pairs, group_first


def run_batch(clouds, pairs, max_range=2.5, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = str(v)
    try:
        lb = pkg.RegistrationBatch(params={"registration_method": "FAST_GICP"})
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ids = {f: lb.addCloud(c) for f, c in clouds.items()}
    res = lb.align([(ids[s], ids[t], g) for s, t, g in pairs], True, max_range)
    rounds = lb.lastRounds()
    lb.close()
    return res, rounds


def same(a, b):
    return np.array_equal(a["T"], b["T"]) and a["converged"] == b["converged"] and a["iterations"] == b["iterations"] and \
        (a["fitness"] == b["fitness"] or (np.isnan(a["fitness"]) and np.isnan(b["fitness"])))


def test_batch_equals_each_pair_alone(synth):
    clouds, pairs, group_first = make_groups(synth)
    res, rounds = run_batch(clouds, pairs)
    assert rounds[0] >= 3 and rounds[1] >= rounds[0]
    # (1) the reference-shaped sequential path: b2r_loop_matching per group (one handle, one candidate after another)
    reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    det = pkg.LoopDetector(reg, fitness_score_max_range=2.5, fitness_score_thresh=0.5)
    for g in range(len(group_first) - 1):
        ps = pairs[group_first[g]: group_first[g + 1]]
        best, rs = det.matching([clouds[s] for s, _, _ in ps], clouds[ps[0][1]], [gg for _, _, gg in ps])
        for r, q in zip(rs, res[group_first[g]: group_first[g + 1]]):
            assert same(r, q)
        assert best == pkg.loop_argmin([_as_result(q) for q in res[group_first[g]: group_first[g + 1]]], 0.5)
    # (2) plain setInputTarget / setInputSource / align on a handle: pose, flag and iteration count bitwise; the fitness of the
    # handle path comes from another kernel (k_fitness, different reduction tree): equal to rounding
    for (s, t, g), q in zip(pairs[:5], res[:5]):
        reg.setInputTarget(clouds[t])
        reg.setInputSource(clouds[s])
        reg.align(g)
        assert np.array_equal(reg.getFinalTransformation(), q["T"]) and reg.hasConverged() == q["converged"] and reg.nr_iterations == q["iterations"]
        f = reg.getFitnessScore(2.5)
        assert abs(f - q["fitness"]) <= 1e-12 * abs(f)
    reg.close()
    # (3) every pair as a batch of one
    for i in (0, 5, 11):
        r1, _ = run_batch(clouds, [pairs[i]])
        assert same(r1[0], res[i])


def _as_result(d):
    from hdl_graph_slam_b200 import _capi
    r = _capi.Result()
    r.fitness, r.converged, r.iterations = d["fitness"], int(d["converged"]), d["iterations"]
    return r


@pytest.mark.parametrize("env", [{"B2R_BATCH_COPIES": 1}, {"B2R_BATCH_COPIES": 2}, {"B2R_BATCH_COPIES": 4}, {"B2R_BATCH_CHUNK": 5}])
def test_batch_invariant_to_lanes_per_query_and_chunking(synth, env):
    clouds, pairs, _ = make_groups(synth, targets=(0, 10), cands=(1, 251, 2))
    ref, _ = run_batch(clouds, pairs)
    got, _ = run_batch(clouds, pairs, env=env)
    for a, b in zip(ref, got):
        assert same(a, b)


def test_batch_matches_oracle(synth, oracle):
    clouds, pairs, _ = make_groups(synth, targets=(0,), cands=(1, 251, 3))
    res, _ = run_batch(clouds, pairs)
    for (s, t, g), q in zip(pairs, res):
        o = oracle.gicp_align(clouds[s], clouds[t], g)
        assert q["converged"] == o["converged"] and q["iterations"] == o["iterations"]
        assert trans_err(q["T"], o["T"]) < 1e-6 and rot_err(q["T"], o["T"]) < 1e-6
        score, _, _ = oracle.fitness(clouds[t], clouds[s], o["T"], 2.5)
        assert abs(q["fitness"] - score) <= 1e-6 * score


def test_batch_mixed_sizes_empty_clouds_and_recycling(synth):
    big = synth.scan("vlp16", frame=0, stride=8)
    small = synth.scan("vlp16_16k", frame=1, stride=8)
    tiny = synth.scan("vlp16_16k", frame=2, stride=8)[:700]
    empty = np.zeros((0, 8), np.float32)
    lb = pkg.RegistrationBatch(params={"registration_method": "FAST_GICP"})
    ib, is_, it, ie = lb.addCloud(big), lb.addCloud(small), lb.addCloud(tiny), lb.addCloud(empty)
    I = np.eye(4, dtype=np.float32)
    g = perturb(3, 0.2, 1.0).astype(np.float32)
    res = lb.align([(is_, ib, I), (ie, ib, g), (it, ib, I), (is_, ie, g), (ib, is_, I)], True, 2.5)
    # empty source / empty target: pcl::Registration::initCompute fails silently -> not converged, final transformation = guess
    for k in (1, 3):
        assert not res[k]["converged"] and res[k]["iterations"] == 0 and np.allclose(res[k]["T"], g)
    alone = []
    reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    for (s, t) in ((small, big), (tiny, big), (big, small)):
        reg.setInputTarget(t)
        reg.setInputSource(s)
        reg.align(I)
        alone.append((reg.getFinalTransformation(), reg.hasConverged(), reg.nr_iterations))
    reg.close()
    for k, a in zip((0, 2, 4), alone):
        assert np.array_equal(res[k]["T"], a[0]) and res[k]["converged"] == a[1] and res[k]["iterations"] == a[2]
    # remove + add: the recycled cloud object (device buffers reused, other size) must give the same answer as a fresh one
    assert lb.cloudCount() == 4
    lb.removeCloud(ib)
    lb.removeCloud(it)
    i2 = lb.addCloud(small)     # lands in a recycled slot that held a different cloud
    i3 = lb.addCloud(big)
    assert lb.cloudCount() == 4
    again = lb.align([(i2, i3, I)], True, 2.5)
    assert same(again[0], res[0])
    with pytest.raises(pkg.B2RError):
        lb.align([(ib if ib not in (i2, i3) else 99, 77, I)])
    lb.close()


def test_loop_detect_single_rank_equals_align_plus_argmin(synth):
    clouds, pairs, group_first = make_groups(synth, targets=(0, 10), cands=(1, 251, 2))
    lb = pkg.RegistrationBatch(params={"registration_method": "FAST_GICP"})
    ids = {f: lb.addCloud(c) for f, c in clouds.items()}
    pl = [(ids[s], ids[t], g) for s, t, g in pairs]
    res = lb.align(pl, True, 2.5)
    best, allres = lb.loopDetect(pl, group_first, 2.5, 0.5)
    for a, b in zip(res, allres):
        assert same(a, b)
    for g in range(2):
        assert best[g] == pkg.loop_argmin([_as_result(q) for q in res[group_first[g]: group_first[g + 1]]], 0.5)
    assert any(b >= 0 for b in best)
    lb.close()


def test_calc_fitness_score_batched_on_cached_keyframes(synth, oracle):
    """InformationMatrixCalculator::calc_fitness_score(cloud1, cloud2, relpose, max_range)
    (src/hdl_graph_slam/information_matrix_calculator.cpp:49-80) for several edges at once, on keyframe clouds registered once:
    kd-tree on cloud1, cloud2 transformed by relpose.cast<float>(), mean squared NN distance over d2 <= max_range."""
    frames = {f: synth.scan("vlp16_16k", frame=f, stride=8) for f in (0, 1, 2, 30)}
    lb = pkg.RegistrationBatch(params={"registration_method": "FAST_GICP"})
    ids = {f: lb.addCloud(c) for f, c in frames.items()}
    edges = []
    for (a, b_) in ((0, 1), (1, 2), (0, 2), (2, 0), (0, 30)):
        rel = (np.linalg.inv(synth.pose_matrix(a)) @ synth.pose_matrix(b_) @ perturb(a * 7 + b_, 0.05, 0.5)).astype(np.float32)
        edges.append((a, b_, rel))
    for max_range in (np.finfo(np.float64).max, 2.0, 0.01):
        got = lb.calcFitnessScore([(ids[a], ids[b_], rel) for a, b_, rel in edges], max_range)
        for (a, b_, rel), g in zip(edges, got):
            want, used, _ = oracle.fitness(frames[a], frames[b_], rel, max_range)
            assert (g == want == np.finfo(np.float64).max) if used == 0 else abs(g - want) <= 1e-12 * want
    # the clouds stay usable for registrations afterwards (covariances are built on demand), and an align leaves the scores unchanged
    again = lb.calcFitnessScore([(ids[0], ids[1], edges[0][2])], 2.0)
    lb.align([(ids[1], ids[0], edges[0][2])], True, 2.0)
    assert lb.calcFitnessScore([(ids[0], ids[1], edges[0][2])], 2.0)[0] == again[0]
    lb.close()


def test_detect_plan_batch_replay_equals_the_sequential_walk(synth):
    """LoopDetector::detect (loop_detector.hpp:57-68) as plan -> b2r_batch_loop_detect -> replay (LoopClosureGate.detect) against the reference's
    sequential walk: find_candidates -> matching (b2r_loop_matching on one handle) -> last_edge_accum_distance update, new keyframe by new keyframe.
    The second new keyframe lies 2 m after the first: it is planned speculatively and dropped by the replay once the first registers its loop."""
    from test_loop_gate import ref_find_candidates
    sensor = "vlp16_16k"
    kf_frames = [0, 2, 10, 12, 20, 120]
    new_frames = [251, 253, 261, 271]
    kfs = [(float(f), synth.pose_matrix(f)) for f in kf_frames]
    news = [(float(f), synth.pose_matrix(f) @ perturb(300 + f, 0.2, 1.0)) for f in new_frames]  # graph estimates drift a little
    clouds = {f: synth.scan(sensor, frame=f, stride=8) for f in kf_frames + new_frames}
    p = dict(distance_thresh=5.0, accum_distance_thresh=8.0, min_edge_interval=5.0)
    gate = pkg.LoopClosureGate(fitness_score_max_range=2.5)
    # (1) the reference's walk on the pcl::Registration-shaped handle
    reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    det = pkg.LoopDetector(reg, fitness_score_max_range=2.5, fitness_score_thresh=0.5)
    last, want = 0.0, []
    for gi, nk in enumerate(news):
        cands = ref_find_candidates(p, kfs, nk, last)
        if not cands:
            continue
        guesses = [gate.guess(nk[1], kfs[c][1]) for c in cands]
        best, results = det.matching([clouds[kf_frames[c]] for c in cands], clouds[new_frames[gi]], guesses)
        if best >= 0:
            want.append((gi, cands[best], results[best]["T"]))
            last = nk[0]
    reg.close()
    assert len(want) >= 2 and all(g != 1 for g, _, _ in want), "the 2 m keyframe must be gated by the loop registered just before it"
    # (2) plan -> batch -> replay
    lb = pkg.RegistrationBatch(params={"registration_method": "FAST_GICP"})
    kid = [lb.addCloud(clouds[f]) for f in kf_frames]
    nid = [lb.addCloud(clouds[f]) for f in new_frames]
    cand, _, gf = gate.plan(kfs, news)
    assert gf[2] - gf[1] >= 1, "the gated keyframe has candidates in the plan (it is matched speculatively)"
    got = gate.detect(lb, kfs, kid, news, nid)
    lb.close()
    assert [(g, c) for g, c, _ in got] == [(g, c) for g, c, _ in want]
    for (_, _, Ta), (_, _, Tb) in zip(got, want):
        assert np.array_equal(Ta, Tb)  # batch == single handle, bitwise
    assert gate.last_edge_accum_distance == last
