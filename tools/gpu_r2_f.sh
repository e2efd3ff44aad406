#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python tools/prof_batch.py 1 1 > $O/sanitizer_batch.log 2>&1
tail -4 $O/sanitizer_batch.log
timeout 1700 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -15 $O/pytest_gpu.txt
timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --workload loop_batch > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
for v in mb10 mb12; do B2R_LIB=$PWD/hdl_graph_slam_b200/_lib/alt/libb200reg_$v.so timeout 600 python bench.py --steps 100 --warmup 5 --no-anchor --no-profile --cpu-sample 0 > $O/bench_n1_$v.json 2> $O/bench_n1_$v.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_odo.csv python bench.py --steps 12 --warmup 3 --no-profile --no-anchor --cpu-sample 0 > $O/ncu_odo.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_batch.csv python tools/prof_batch.py 8 1 > $O/ncu_batch.log 2>&1
