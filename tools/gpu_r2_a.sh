#!/bin/bash
# round-2 first GPU pass: parity suite, smoke, odometry bench, loop-batch bench
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/gpu.txt 2>&1
nproc >> gpurun_out/r2a/gpu.txt
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2a/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a/smoke.txt 2>&1
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2a/bench_n1.json 2> gpurun_out/r2a/bench_n1.err
timeout 600 python bench.py --workload loop_batch > gpurun_out/r2a/bench_loop_n1.json 2> gpurun_out/r2a/bench_loop_n1.err
for c in 2 4; do B2R_BATCH_COPIES=$c timeout 300 python bench.py --workload loop_batch --no-profile --steps 3 > gpurun_out/r2a/bench_loop_n1_c$c.json 2> gpurun_out/r2a/bench_loop_n1_c$c.err; done
tail -3 gpurun_out/r2a/pytest_gpu.txt
