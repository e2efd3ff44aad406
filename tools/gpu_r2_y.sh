#!/bin/bash
# 1 GPU: accumulate at 4 blocks/SM (product), A/B of 5 blocks/SM and of the k-NN kernel at 5 blocks/SM; ncu --set full of the round-2 final kernels
O=gpurun_out/r2y; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so hdl_graph_slam_b200/_lib/alt/*.so > $O/lib.md5
A=$PWD/hdl_graph_slam_b200/_lib/alt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 env B2R_LIB=$A/libb200reg_knn5.so python bench.py --steps 200 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_n1_knn5.json 2> $O/bench_n1_knn5.err
timeout 400 python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 400 env B2R_LIB=$A/libb200reg_acc5.so python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1_acc5.json 2> $O/bench_loop_n1_acc5.err
timeout 400 env B2R_LIB=$A/libb200reg_knn5.so python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1_knn5.json 2> $O/bench_loop_n1_knn5.err
for f in bench_n1 bench_n1_knn5 bench_loop_n1 bench_loop_n1_acc5 bench_loop_n1_knn5; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("roofline") or {}).get("avg_launch_us"), d["config"].get("strict_chain_value"), d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_search|k_pair_accumulate|k_knn_cov_reg_batch" -s 3 -c 5 -o $O/prof_batch python tools/prof_batch.py 8 1 > $O/ncu_full_batch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_search|k_pair_accumulate|k_pair_lm|k_knn_cov_reg|k_bvh_build" -s 12 -c 8 -o $O/prof_one python tools/prof_one.py > $O/ncu_full_one.log 2>&1
