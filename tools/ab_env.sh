#!/bin/bash
# tests + standalone preprocessing times + odometry bench on one box, then the instrumented k-NN variant
gicp() { python bench.py --steps ${STEPS:-100} --warmup 5 --cpu-sample 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('gicp', round(d['value'],1), round(d['e2e']['value'],1), d['kernel_ms_in_timed_region'], d['config']['mean_iterations'])"; }
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/knn_time.py
gicp; gicp
cp hdl_graph_slam_b200/_lib/alt/libb200reg_knnprof.so hdl_graph_slam_b200/_lib/libb200reg.so; python tools/knn_time.py 3
