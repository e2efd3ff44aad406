#!/bin/bash
# 1 GPU: suite on the current build; the driver's own bench form (20 steps: clock sampling with hold steps); stream-priority A/B on the odometry
# chain; NDT chunk-order A/B; accumulate occupancy A/B on the loop batch
O=gpurun_out/r2x; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so hdl_graph_slam_b200/_lib/alt/*.so > $O/lib.md5
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
A=$PWD/hdl_graph_slam_b200/_lib/alt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_form.json 2> $O/bench_driver_form.err
timeout 600 python bench.py --steps 200 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 env B2R_NO_PRIORITY=1 python bench.py --steps 200 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_n1_noprio.json 2> $O/bench_n1_noprio.err
timeout 600 env B2R_PREFETCH_PRIORITY=1 python bench.py --steps 200 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_n1_prefprio.json 2> $O/bench_n1_prefprio.err
timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
timeout 600 env B2R_LIB=$A/libb200reg_ndtstride.so python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1_stride.json 2> $O/bench_ndt_n1_stride.err
timeout 400 python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 400 env B2R_LIB=$A/libb200reg_acc4.so python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1_acc4.json 2> $O/bench_loop_n1_acc4.err
for f in bench_driver_form bench_n1 bench_n1_noprio bench_n1_prefprio bench_ndt_n1 bench_ndt_n1_stride bench_loop_n1 bench_loop_n1_acc4; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("roofline") or {}).get("avg_launch_us"), d["config"].get("strict_chain_value"), d.get("kernel_ms_in_timed_region"), {k: v for k, v in (d.get("clocks") or {}).items() if k != "window"})
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
