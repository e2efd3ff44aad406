#!/bin/bash
# 1 GPU: transposed partial sums with a coalesced final reduction (NDT pass, k_pair_lm) — suite, benches, NDT split A/B, ncu of the NDT pass
O=gpurun_out/r2t; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so > $O/lib.md5
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
timeout 600 env B2R_NDT_EVEN=1 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1_even.json 2> $O/bench_ndt_n1_even.err
timeout 600 python bench.py --steps 200 --warmup 5 --cpu-sample 0 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 400 python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 900 python bench.py --workload kitti_pipeline --steps 100 --warmup 5 --cpu-sample 0 > $O/bench_kitti.json 2> $O/bench_kitti.err
for f in bench_ndt_n1 bench_ndt_n1_even bench_n1 bench_loop_n1 bench_kitti; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("roofline") or {}).get("avg_launch_us"), d["config"].get("strict_chain_value"), d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ndt_derivatives" -s 3 -c 1 -o $O/prof_ndt python tools/prof_ndt.py > $O/ncu_full_ndt.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_odo.csv python bench.py --steps 4 --warmup 3 --no-profile --cpu-sample 0 --no-anchor > $O/ncu_odo.log 2>&1
