#!/bin/bash
# 1 GPU: residency cap of the prefetch stream's k-NN kernel (B2R_PREFETCH_KNN_BLOCKS = blocks per SM) on the odometry chain
O=gpurun_out/r2aa; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so > $O/lib.md5
for c in 0 1 2 3; do
  timeout 600 env B2R_PREFETCH_KNN_BLOCKS=$c python bench.py --steps 200 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_n1_cap$c.json 2> $O/bench_n1_cap$c.err
done
timeout 900 env B2R_PREFETCH_KNN_BLOCKS=2 python bench.py --workload kitti_pipeline --steps 100 --warmup 5 --cpu-sample 0 > $O/bench_kitti_cap2.json 2> $O/bench_kitti_cap2.err
for f in bench_n1_cap0 bench_n1_cap1 bench_n1_cap2 bench_n1_cap3 bench_kitti_cap2; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("roofline") or {}).get("avg_launch_us"), d["config"].get("strict_chain_value"), d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
timeout 600 env B2R_PREFETCH_KNN_BLOCKS=2 python -m pytest tests/test_callers_gpu.py tests/test_gicp_gpu.py -m gpu -q -x > $O/pytest_cap2.txt 2>&1; tail -2 $O/pytest_cap2.txt
