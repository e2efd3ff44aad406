#!/bin/bash
# 1 GPU: suite + every bench workload on the build with the minimum-first 1-NN tile sweep; ncu launch lists + full capture of the batched round
O=gpurun_out/r2n; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so > $O/lib.md5
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -4 $O/pytest_gpu.txt
timeout 400 python bench.py --workload loop_batch > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 600 python bench.py --steps 200 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
timeout 600 python bench.py --workload voxelgrid --steps 200 --warmup 5 > $O/bench_voxelgrid.json 2> $O/bench_voxelgrid.err
timeout 900 python bench.py --workload kitti_pipeline --steps 100 --warmup 5 > $O/bench_kitti.json 2> $O/bench_kitti.err
for f in bench_loop_n1 bench_n1 bench_ndt_n1 bench_voxelgrid bench_kitti; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("roofline") or {}).get("avg_launch_us"), (d.get("cpu_baseline") or {}).get("value"), d["config"].get("strict_chain_value"), d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_batch.csv python bench.py --workload loop_batch --steps 1 --warmup 1 --pairs 64 --no-profile --cpu-sample 0 > $O/ncu_batch.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_odo.csv python bench.py --steps 4 --warmup 3 --no-profile --cpu-sample 0 --no-anchor > $O/ncu_odo.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_search|k_pair_accumulate" -s 2 -c 4 -o $O/prof_batch python tools/prof_batch.py 8 1 > $O/ncu_full_batch.log 2>&1
