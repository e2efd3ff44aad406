#!/bin/bash
# 1 GPU: packed-key 1-NN visitor in k_pair_search — suite, benches; micro-benchmark with the FFMA-immediate variant
O=gpurun_out/r2v; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so > $O/lib.md5
timeout 120 tools/microbench/tile_f32x2 40 > $O/microbench_tile.txt 2>&1; cat $O/microbench_tile.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 5 --cpu-sample 0 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 400 python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 900 python bench.py --workload kitti_pipeline --steps 100 --warmup 5 --cpu-sample 0 > $O/bench_kitti.json 2> $O/bench_kitti.err
for f in bench_n1 bench_loop_n1 bench_kitti; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("roofline") or {}).get("avg_launch_us"), d["config"].get("strict_chain_value"), d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
