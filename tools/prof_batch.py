"""tiny driver for ncu captures of the batched path: N candidate pairs (64k-point VLP-16) through b2r_batch_align, one pass"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import synth, batch

n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 8
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
groups, guesses, group_first = batch.loop_workload(n_groups, "vlp16")
frames = sorted({f for tf, sfs in groups for f in [tf] + sfs})
clouds = {f: synth.scan("vlp16", frame=f, stride=8) for f in frames}
lb = pkg.RegistrationBatch(params={"registration_method": "FAST_GICP"})
for _ in range(passes):
    ids = {f: lb.addCloud(c) for f, c in clouds.items()}
    pairs = [(ids[sf], ids[tf], guesses[group_first[g] + c]) for g, (tf, sfs) in enumerate(groups) for c, sf in enumerate(sfs)]
    res = lb.align(pairs, True, 2.5)
    for i in ids.values():
        lb.removeCloud(i)
print("pairs", len(pairs), "clouds", len(frames), "rounds", lb.lastRounds(), "converged", sum(r["converged"] for r in res),
      "iters", sum(r["iterations"] for r in res) / len(res))
lb.close()
