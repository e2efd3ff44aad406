#!/bin/bash
# 1 GPU: dynamic lane regrouping of sparse tile visits in the batched (1 lane per query) packed-key search — suite, loop batch A/B, odometry
O=gpurun_out/r2ad; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so hdl_graph_slam_b200/_lib/alt/*.so > $O/lib.md5
A=$PWD/hdl_graph_slam_b200/_lib/alt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
timeout 400 python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 400 env B2R_LIB=$A/libb200reg_noregroup.so python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1_noregroup.json 2> $O/bench_loop_n1_noregroup.err
timeout 600 python bench.py --steps 200 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_n1.json 2> $O/bench_n1.err
for f in bench_loop_n1 bench_loop_n1_noregroup bench_n1; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
