#!/bin/bash
# one gpurun call: parity tests, smoke, all bench workloads + reference arm, ncu launch list and one ncu --set full capture.
# Outputs land in gpurun_out/final/.
O=gpurun_out/final
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.txt
echo "=== bench gicp odometry"; python bench.py 2> $O/bench_err.log | tee $O/bench_n1.json | cut -c1-300
echo "=== bench reference arm"; python bench.py --impl reference --steps 6 --warmup 1 2>> $O/bench_err.log | tee $O/bench_ref.json | cut -c1-300
echo "=== bench ndt odometry"; python bench.py --workload ndt_odometry_hdl32e_128k --steps 30 --warmup 3 --cpu-sample 2 2>> $O/bench_err.log | tee $O/bench_ndt_n1.json | cut -c1-300
echo "=== bench loop batch"; python bench.py --workload loop_batch --pairs ${PAIRS:-256} 2>> $O/bench_err.log | tee $O/bench_loop_n1.json | cut -c1-300
tail -3 $O/bench_err.log
echo "=== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 12 --warmup 3 --cpu-sample 0 --no-profile > $O/ncu_bench.log 2>&1
python tools/agg_launches.py $O/launches.csv | tee $O/launches.md
[ -z "$NCU_FULL" ] && exit 0
echo "=== ncu full"
ncu --set full --clock-control none --import-source on -k regex:"k_gicp_correspond|k_knn_cov_reg|k_gicp_accumulate" -c 6 -o $O/prof_final python tools/prof_one.py > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
ls -la $O
