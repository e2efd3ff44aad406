#!/bin/bash
# tests + standalone preprocessing times + odometry benches on one box, then the instrumented variant (per-warp profiles)
gicp() { python bench.py --steps ${STEPS:-100} --warmup 5 --cpu-sample 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('gicp', round(d['value'],1), round(d['e2e']['value'],1), d['kernel_ms_in_timed_region'], d['config']['mean_iterations'])"; }
ndt() { python bench.py --workload ndt_odometry_hdl32e_128k --steps 30 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ndt', round(d['value'],1), d['kernel_ms_in_timed_region'])"; }
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/knn_time.py
gicp; gicp; ndt
cp hdl_graph_slam_b200/_lib/alt/libb200reg_knnprof.so hdl_graph_slam_b200/_lib/libb200reg.so; python tools/prof_one.py > /dev/null 2>&1; ls gpurun_out | grep prof_
