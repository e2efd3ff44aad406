#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -25 $O/pytest_gpu.txt
timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --workload loop_batch > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_ndt.csv python bench.py --workload ndt_odometry_hdl32e_128k --steps 4 --warmup 3 --no-profile --cpu-sample 0 --no-anchor > $O/ncu_ndt.log 2>&1
