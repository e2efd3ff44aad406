#!/bin/bash
# tests + all three bench workloads on one GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "=== bench gicp odometry"; python bench.py --steps ${STEPS:-100} --warmup 5 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | cut -c1-200
echo "=== bench ndt odometry"; python bench.py --workload ndt_odometry_hdl32e_128k --steps 30 --warmup 3 --cpu-sample 2 2> gpurun_out/bench_ndt_err.log | tee gpurun_out/bench_ndt_n1.json | cut -c1-1500
tail -3 gpurun_out/bench_ndt_err.log
echo "=== bench loop batch"; python bench.py --workload loop_batch --pairs ${PAIRS:-64} 2> gpurun_out/bench_loop_err.log | tee gpurun_out/bench_loop_n1.json | cut -c1-1200
tail -3 gpurun_out/bench_loop_err.log
