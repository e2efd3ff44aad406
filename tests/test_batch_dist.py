"""CPU-only (gloo, world_size 2): the sharding map, the single fixed-size all-gather and the per-group argmin of the loop-closure
batch give the same answer as a single process; the Python mirror of the layout equals the library's own b2r_shard_range /
b2r_loop_argmin (host-only entry points: they run without a GPU); the tie rule follows loop_detector.hpp:147."""
import os
import socket
import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import batch, _capi


def fake_records(group_first, g0, g1):
    """deterministic stand-in for the GPU work: a record that depends only on (group, candidate)"""
    recs = []
    for g in range(g0, g1):
        for c in range(group_first[g + 1] - group_first[g]):
            r = np.zeros((), batch.RECORD_DTYPE)
            r["T"] = np.arange(16, dtype=np.float32) + 100 * g + c
            r["fitness"] = ((g * 7919 + c * 104729) % 1000) / 1000.0 + 0.01
            r["converged"] = 0 if (g + c) % 5 == 0 else 1
            r["iterations"] = 3 + c
            recs.append(r)
    return np.array(recs, batch.RECORD_DTYPE) if recs else np.zeros(0, batch.RECORD_DTYPE)


def _worker(rank, world, port, group_first, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M, ranges = batch.layout(group_first, world)
    g0, g1 = batch.shard_range(len(group_first) - 1, world, rank)
    local = fake_records(group_first, g0, g1)
    assert len(local) == ranges[rank][1] - ranges[rank][0]
    gathered = batch.gather_records_torch(local, M, world, None)
    records = batch.unpack_gathered(gathered, group_first, world)
    best = batch.argmin_per_group(records, group_first, 0.5)
    q.put((rank, best, records.tobytes()))
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    sizes = [8, 3, 0, 5, 8, 1, 7]
    group_first = [0] + list(np.cumsum(sizes))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, group_first, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rec1 = fake_records(group_first, 0, len(sizes))
    best1 = batch.argmin_per_group(rec1, group_first, 0.5)
    for rank, best, flat in outs:
        assert best == best1
        assert flat == rec1.tobytes()  # every rank holds every pair's record, bit-identical to the 1-process run


def test_layout_mirror_equals_library():
    for n_groups in (0, 1, 7, 8, 256, 257):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for rank in range(world):
                assert batch.shard_range(n_groups, world, rank) == pkg.shard_range(n_groups, world, rank)
                covered.append(batch.shard_range(n_groups, world, rank))
            assert covered[0][0] == 0 and covered[-1][1] == n_groups and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    M, ranges = batch.layout([0, 2, 4, 6], 2)
    assert M == 4 and ranges == [(0, 2), (2, 6)]


def test_argmin_tie_rule_library_and_mirror():
    def both(fit, conv, thresh=0.5):
        recs = np.zeros(len(fit), batch.RECORD_DTYPE)
        recs["fitness"], recs["converged"] = fit, conv
        rs = []
        for f, c in zip(fit, conv):
            r = _capi.Result()
            r.fitness, r.converged = f, c
            rs.append(r)
        a = batch.argmin_per_group(recs, [0, len(fit)], thresh)[0]
        b = pkg.loop_argmin(rs, thresh)
        assert a == b
        return a
    # `score > best_score -> continue`: an equal score REPLACES the earlier candidate (loop_detector.hpp:147)
    assert both([0.2, 0.2, 0.3], [1, 1, 1]) == 1
    assert both([0.2, 0.2, 0.3], [1, 0, 1]) == 0
    assert both([0.7, 0.6, 0.9], [1, 1, 1]) == -1  # best above fitness_score_thresh: "loop not found"
    assert both([], []) == -1
    assert both([0.1, 0.2], [0, 0]) == -1


def test_information_matrix_from_fitness_matches_reference_formula():
    """information_matrix_calculator.cpp:25-47 + weight() (information_matrix_calculator.hpp:39-42), restated with numpy"""
    def ref(f, a=20.0, thr=0.5, mnx=0.1, mxx=5.0, mnq=0.05, mxq=0.2):
        def w(min_y, max_y):
            y = (1.0 - np.exp(-a * f)) / (1.0 - np.exp(-a * thr))
            return np.float32(min_y + (max_y - min_y) * y)
        return np.array([1.0 / np.float64(w(mnx ** 2, mxx ** 2))] * 3 + [1.0 / np.float64(w(mnq ** 2, mxq ** 2))] * 3)
    for f in (0.0, 0.01, 0.2, 0.5, 3.0):
        assert np.allclose(pkg.information_from_fitness(f), ref(f), rtol=1e-15, atol=0)
    assert np.allclose(pkg.information_from_fitness(0.3, fitness_score_thresh=2.5, var_gain_a=10.0), ref(0.3, a=10.0, thr=2.5), rtol=1e-15)
    assert np.allclose(pkg.information_from_fitness(0.3, use_const_inf_matrix=1), [2.0, 2.0, 2.0, 10.0, 10.0, 10.0])


# ---- LoopDetector::detect across two ranks: plan on every rank, each rank "matches" its shard, one gather, sequential replay
def _detect_inputs():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_loop_gate import trajectory
    rng = np.random.default_rng(21)
    kfs = trajectory(rng, 60)
    new = []
    for k in (3, 4, 9, 17, 18, 30, 41, 42, 55):
        T = kfs[k][1].copy()
        T[:3, 3] += rng.normal(size=3) * 0.4
        new.append((kfs[k][0] + 120.0, T))
    return kfs, new


def _fake_match(group_first, g0, g1, cand):
    """a record per (new keyframe, candidate) that depends only on the pair: some groups find a loop, some do not"""
    recs = []
    for g in range(g0, g1):
        for j in range(group_first[g], group_first[g + 1]):
            r = np.zeros((), batch.RECORD_DTYPE)
            r["T"] = np.arange(16, dtype=np.float32) + 10 * g + cand[j]
            r["fitness"] = 0.9 if g % 4 == 1 else 0.05 + 0.01 * ((cand[j] * 31 + g) % 17)
            r["converged"] = 1
            r["iterations"] = 4
            recs.append(r)
    return np.array(recs, batch.RECORD_DTYPE) if recs else np.zeros(0, batch.RECORD_DTYPE)


def _detect_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kfs, new = _detect_inputs()
    gate = pkg.LoopClosureGate()
    gate.last_edge_accum_distance = 100.0
    cand, _, gf = gate.plan(kfs, new)          # every rank computes the same plan from the same keyframe states
    M, ranges = batch.layout(gf, world)
    g0, g1 = batch.shard_range(len(new), world, rank)
    local = _fake_match(gf, g0, g1, cand)       # stands in for the rank's share of b2r_batch_loop_detect
    records = batch.unpack_gathered(batch.gather_records_torch(local, M, world, None), gf, world)
    best = batch.argmin_per_group(records, gf, gate.params.fitness_score_thresh)
    accepted = gate.replay(new, gf, best, 100.0)
    q.put((rank, [(g, cand[gf[g] + a]) for g, a in enumerate(accepted) if a >= 0], gate.last_edge_accum_distance))
    dist.destroy_process_group()


def test_two_rank_detect_walk_equals_the_sequential_reference_walk():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_loop_gate import ref_find_candidates
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_detect_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the reference's sequential walk (loop_detector.hpp:57-68) with the same fake matching
    kfs, new = _detect_inputs()
    p = dict(distance_thresh=5.0, accum_distance_thresh=8.0, min_edge_interval=5.0)
    last, want = 100.0, []
    for g, nk in enumerate(new):
        cands = ref_find_candidates(p, kfs, nk, last)
        if not cands:
            continue
        gf = [0, len(cands)]
        rec = _fake_match([0] * g + gf, g, g + 1, cands)
        b = batch.argmin_per_group(rec, gf, 0.5)[0]
        if b >= 0:
            want.append((g, cands[b]))
            last = nk[0]
    assert len(want) >= 3
    for rank, loops, last_edge in outs:
        assert loops == want and last_edge == last
