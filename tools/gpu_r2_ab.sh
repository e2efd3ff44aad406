#!/bin/bash
# 1 GPU: tile / cooperative threshold of the packed-key search (interested lanes at which a leaf visit switches to the all-pairs tile) on the loop batch and the odometry chain
O=gpurun_out/r2ab; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so hdl_graph_slam_b200/_lib/alt/*.so > $O/lib.md5
A=$PWD/hdl_graph_slam_b200/_lib/alt
timeout 400 python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
for v in tl5 tl12 tl16; do
  timeout 400 env B2R_LIB=$A/libb200reg_$v.so python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1_$v.json 2> $O/bench_loop_n1_$v.err
done
timeout 600 env B2R_LIB=$A/libb200reg_tl12.so python bench.py --steps 200 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_n1_tl12.json 2> $O/bench_n1_tl12.err
for f in bench_loop_n1 bench_loop_n1_tl5 bench_loop_n1_tl12 bench_loop_n1_tl16 bench_n1_tl12; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
