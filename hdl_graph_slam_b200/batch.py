"""Loop-closure candidate batch across GPUs (BASELINE.json configs[3]).

The reference validates loop candidates one after another inside LoopDetector::matching
(/root/reference/include/hdl_graph_slam/loop_detector.hpp:117-171): one shared target (the new keyframe, :122) and C
independent align() + getFitnessScore() calls (:135-154).  The calls only couple through the running `best_score`, i.e. a
post-hoc argmin, so the batch shards naturally:

  * candidate groups (all candidates of one new keyframe share a target) are dealt round-robin to ranks, so each target's
    grid / covariances / voxel map is built once per group on one GPU;
  * clouds reach a GPU only over PCIe (H2D at set_target / set_source); nothing point-sized ever crosses NVLink;
  * ONE collective ends the batch: an all-gather of fixed 80-byte records {T[16] f32, fitness f64, converged i32,
    iterations i32} (== b2r_result) over NCCL (torch.distributed) — padded to the largest per-rank count;
  * every rank then holds all records and runs the reference's argmin per group on the host (same tie rule as
    loop_detector.hpp:147: a later candidate with an EQUAL score replaces the earlier one).

One process per GPU (torchrun).  Within a GPU, `streams_per_gpu` registration handles are driven by host threads so that
independent registrations overlap on the device (each handle has its own CUDA stream and buffers, as in the reference's
"two handles live at once" threading model).
"""
import ctypes as C
import threading
import time
import numpy as np

RECORD_BYTES = 80
RECORD_DTYPE = np.dtype([("T", np.float32, (16,)), ("fitness", np.float64), ("converged", np.int32), ("iterations", np.int32)])
assert RECORD_DTYPE.itemsize == RECORD_BYTES


def shard_groups(n_groups, world):
    """group g -> rank g % world (every rank can recompute the whole assignment)"""
    return [[g for g in range(n_groups) if g % world == r] for r in range(world)]


def layout(group_sizes, world):
    """Deterministic placement of every (group, candidate) in the gathered buffer: returns (per_rank_counts, slot[(g, c)] = (rank, local index))"""
    owners = shard_groups(len(group_sizes), world)
    counts, slot = [], {}
    for r, gs in enumerate(owners):
        k = 0
        for g in gs:
            for c in range(group_sizes[g]):
                slot[(g, c)] = (r, k)
                k += 1
        counts.append(k)
    return counts, slot


def argmin_per_group(records, group_sizes, slot, fitness_score_thresh):
    """LoopDetector::matching's selection (loop_detector.hpp:147-163) per group -> best candidate index or -1"""
    best = []
    for g, size in enumerate(group_sizes):
        best_score, best_c = np.finfo(np.float64).max, -1
        for c in range(size):
            r, k = slot[(g, c)]
            rec = records[r][k]
            score = float(rec["fitness"])
            if not rec["converged"] or score > best_score:
                continue
            best_score, best_c = score, c
        if best_score > fitness_score_thresh:
            best_c = -1
        best.append(best_c)
    return best


def gather_records(local_records, counts, rank, world, device=None):
    """ONE all-gather of the fixed-size records (NCCL when `device` is a CUDA device, gloo on CPU). Returns per-rank record arrays."""
    import torch
    import torch.distributed as dist
    pad = max(max(counts), 1)
    buf = np.zeros(pad, RECORD_DTYPE)
    buf[: len(local_records)] = local_records
    t = torch.from_numpy(buf.view(np.uint8).reshape(pad * RECORD_BYTES).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * pad * RECORD_BYTES, dtype=torch.uint8, device=t.device)
    if world > 1:
        dist.all_gather_into_tensor(out, t)
    else:
        out.copy_(t)
    raw = out.cpu().numpy().reshape(world, pad * RECORD_BYTES)
    return [raw[r].view(RECORD_DTYPE)[: counts[r]].copy() for r in range(world)]


class LoopBatch:
    """targets: list of clouds; candidates: list (per target) of lists of (cloud, guess4x4).  Each rank passes the FULL
    description (clouds it does not own may be None)."""

    def __init__(self, params=None, device_id=0, streams_per_gpu=4, fitness_score_max_range=np.finfo(np.float64).max,
                 fitness_score_thresh=0.5):
        from . import registration as R
        self.R = R
        self.params = params or {"registration_method": "FAST_GICP"}
        self.device_id = device_id
        self.n_streams = max(1, streams_per_gpu)
        self.max_range = fitness_score_max_range
        self.thresh = fitness_score_thresh
        self.handles = [R.select_registration_method(self.params, device_id=device_id) for _ in range(self.n_streams)]

    def close(self):
        for h in self.handles:
            h.close()
        self.handles = []

    def _run_group(self, reg, target, cands, out, base):
        reg.setInputTarget(target)                      # loop_detector.hpp:122
        for c, (cloud, guess) in enumerate(cands):      # :135-154
            reg.setInputSource(cloud)
            reg.align(guess)
            score = reg.getFitnessScore(self.max_range)
            rec = out[base + c]
            rec["T"] = np.asarray(reg.getFinalTransformation(), np.float32).T.reshape(-1)  # column-major, as b2r_result
            rec["fitness"] = score
            rec["converged"] = int(reg.hasConverged())
            rec["iterations"] = reg.nr_iterations

    def run_local(self, targets, candidates, my_groups):
        """all groups owned by this rank -> record array in layout order"""
        sizes = [len(candidates[g]) for g in my_groups]
        out = np.zeros(sum(sizes), RECORD_DTYPE)
        bases = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
        work = list(range(len(my_groups)))
        lock = threading.Lock()
        errors = []

        def worker(reg):
            while True:
                with lock:
                    if not work:
                        return
                    j = work.pop(0)
                try:
                    g = my_groups[j]
                    self._run_group(reg, targets[g], candidates[g], out, bases[j])
                except Exception as e:  # noqa: BLE001
                    errors.append(e)
                    return

        threads = [threading.Thread(target=worker, args=(h,)) for h in self.handles]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out

    def run(self, targets, candidates, rank=0, world=1, device=None):
        group_sizes = [len(c) for c in candidates]
        counts, slot = layout(group_sizes, world)
        mine = shard_groups(len(group_sizes), world)[rank]
        local = self.run_local(targets, candidates, mine)
        records = gather_records(local, counts, rank, world, device)
        best = argmin_per_group(records, group_sizes, slot, self.thresh)
        return best, records, slot


# ------------------------------------------------------------------------------------------------ bench leg (configs[3])
def bench_loop_batch(args, rank, world, local_rank):
    import json
    import torch
    import torch.distributed as dist
    from . import synth
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    per_gpu, group = args.pairs, 8
    n_groups_local = max(1, per_gpu // group)
    n_groups = n_groups_local * world
    rng = np.random.default_rng(1234)
    targets, candidates = [None] * n_groups, [[] for _ in range(n_groups)]
    mine = set(shard_groups(n_groups, world)[rank])
    lap = 251
    for g in range(n_groups):
        tf = 3 * g
        for c in range(group):
            sf = tf + lap + c - group // 2  # one lap later, a few metres around the target pose
            dt, da = rng.uniform(-0.5, 0.5, 3), np.deg2rad(rng.uniform(-3, 3))
            if g in mine:
                rel = np.linalg.inv(synth.pose_matrix(tf)) @ synth.pose_matrix(sf)
                P = np.eye(4)
                P[:3, 3] = dt
                P[0, 0], P[0, 1], P[1, 0], P[1, 1] = np.cos(da), -np.sin(da), np.sin(da), np.cos(da)
                guess = (rel @ P).astype(np.float32)
                guess[2, 3] = 0.0  # loop_detector.hpp:142
                candidates[g].append((synth.scan("vlp16", frame=sf), guess))
            else:
                candidates[g].append((None, None))
        if g in mine:
            targets[g] = synth.scan("vlp16", frame=tf)
    n_streams = getattr(args, "streams", 16)
    lb = LoopBatch({"registration_method": "FAST_GICP"}, device_id=local_rank, streams_per_gpu=n_streams, fitness_score_max_range=2.5)
    # warm-up: one group per handle (first-use allocations of every handle happen outside the timed region)
    warm = sorted(mine)[:n_streams]
    lb.run_local(targets, candidates, warm)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # the batch is driven by host threads (one per handle), so a single pass is at the mercy of the host scheduler: time the whole
    # batch `reps` times (identical inputs, every pass redoes all uploads / builds / aligns / the all-gather) and report the median
    reps = 3
    times = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        best, records, slot = lb.run(targets, candidates, rank, world, dev)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        times.append(float(tt.item()))
    t = torch.tensor([sorted(times)[reps // 2]], dtype=torch.float64, device=dev)
    if rank == 0:
        n_pairs = n_groups * group
        conv = sum(int(r["converged"].sum()) for r in records)
        iters = sum(int(r["iterations"].sum()) for r in records)
        print(json.dumps({
            "metric": "registrations/sec", "value": n_pairs / (t.item() * 1e-3), "unit": "registrations/s", "n_gpus": world, "steps": reps, "warmup": 1,
            "ms_per_step": t.item(), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 NN / f64 accumulate",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: loop-closure candidate batch (GICP, 64k-pt VLP-16 pairs, targets shared by 8 candidates)",
                       "pairs_total": n_pairs, "pairs_per_gpu": per_gpu, "collective": "one NCCL all-gather of 80-byte records",
                       "streams_per_gpu": n_streams, "timing": "host clock around the whole batch incl. H2D, max over ranks (host-driven); median of the passes", "pass_ms": [round(x, 2) for x in times]},
            "converged": conv, "mean_iterations": iters / n_pairs, "loops_found": int(sum(1 for b in best if b >= 0)), "groups": n_groups,
        }), flush=True)
    lb.close()
    if world > 1:
        dist.destroy_process_group()
