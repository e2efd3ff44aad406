"""CPU-only (gloo, world_size 2): sharding, the single fixed-size all-gather and the per-group argmin of the loop-closure
batch give the same answer as a single process; the tie rule follows loop_detector.hpp:147."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from hdl_graph_slam_b200 import batch


def fake_records(group_sizes, groups):
    """deterministic stand-in for the GPU work: a record that depends only on (group, candidate)"""
    recs = []
    for g in groups:
        for c in range(group_sizes[g]):
            r = np.zeros((), batch.RECORD_DTYPE)
            r["T"] = np.arange(16, dtype=np.float32) + 100 * g + c
            r["fitness"] = ((g * 7919 + c * 104729) % 1000) / 1000.0 + 0.01
            r["converged"] = 0 if (g + c) % 5 == 0 else 1
            r["iterations"] = 3 + c
            recs.append(r)
    return np.array(recs, batch.RECORD_DTYPE) if recs else np.zeros(0, batch.RECORD_DTYPE)


def _worker(rank, world, port, group_sizes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts, slot = batch.layout(group_sizes, world)
    mine = batch.shard_groups(len(group_sizes), world)[rank]
    local = fake_records(group_sizes, mine)
    records = batch.gather_records(local, counts, rank, world, None)
    best = batch.argmin_per_group(records, group_sizes, slot, 0.5)
    flat = [(g, c, records[slot[(g, c)][0]][slot[(g, c)][1]].tobytes()) for g in range(len(group_sizes)) for c in range(group_sizes[g])]
    q.put((rank, best, flat))
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    group_sizes = [8, 3, 0, 5, 8, 1, 7]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, group_sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    counts1, slot1 = batch.layout(group_sizes, 1)
    rec1 = [fake_records(group_sizes, list(range(len(group_sizes))))]
    best1 = batch.argmin_per_group(rec1, group_sizes, slot1, 0.5)
    flat1 = [(g, c, rec1[0][slot1[(g, c)][1]].tobytes()) for g in range(len(group_sizes)) for c in range(group_sizes[g])]
    for rank, best, flat in outs:
        assert best == best1
        assert flat == flat1  # every rank holds every pair's record, bit-identical to the 1-process run


def test_layout_and_tie_rule():
    counts, slot = batch.layout([2, 2, 2], 2)
    assert counts == [4, 2] and slot[(2, 1)] == (0, 3) and slot[(1, 0)] == (1, 0)
    recs = np.zeros(3, batch.RECORD_DTYPE)
    recs["fitness"] = [0.2, 0.2, 0.3]
    recs["converged"] = [1, 1, 1]
    _, slot = batch.layout([3], 1)
    # `score > best_score -> continue`: an equal score REPLACES the earlier candidate (loop_detector.hpp:147)
    assert batch.argmin_per_group([recs], [3], slot, 0.5) == [1]
    recs["converged"] = [1, 0, 1]
    assert batch.argmin_per_group([recs], [3], slot, 0.5) == [0]
    recs["fitness"] = [0.7, 0.6, 0.9]
    assert batch.argmin_per_group([recs], [3], slot, 0.5) == [-1]  # best above fitness_score_thresh: "loop not found"
    assert batch.argmin_per_group([np.zeros(0, batch.RECORD_DTYPE)], [0], {}, 0.5) == [-1]
