#!/bin/bash
# round-2 second GPU pass: full parity suite, benches, ncu launch lists + full captures of the pair-engine kernels
O=gpurun_out/r2b; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --workload loop_batch > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_odo.csv python bench.py --steps 12 --warmup 3 --no-profile --no-anchor --cpu-sample 0 > $O/ncu_odo.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_batch.csv python tools/prof_batch.py 8 1 > $O/ncu_batch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pair_search|k_pair_accumulate" -s 4 -c 4 -o $O/prof_pair_batch python tools/prof_batch.py 8 1 > $O/ncu_full_batch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pair_search|k_pair_accumulate" -s 4 -c 4 -o $O/prof_pair_single python tools/prof_one.py > $O/ncu_full_single.log 2>&1
tail -3 $O/pytest_gpu.txt
