#!/bin/bash
# 1 GPU: suite on the build with the REDUX / pipelined emit phase of the cluster build; latency-sensitive workloads, and the ballot-ranking variant of the cluster sort
O=gpurun_out/r2m; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -4 $O/pytest_gpu.txt
run() { name=$1; wl=$2; steps=$3; shift 3; timeout 600 env "$@" python bench.py --workload $wl --steps $steps --warmup 5 --cpu-sample 0 --no-anchor > $O/${wl}_$name.json 2> $O/${wl}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/${wl}_$name.json").read().strip().splitlines()[-1]); print("$wl $name", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d["config"].get("strict_chain_value"), (d.get("kernel_ms_in_timed_region") or {}).get("bvh_build"))
except Exception as e: print("$wl $name ERR", e); print(open("$O/${wl}_$name.err").read()[-600:])
PY
}
run base kitti_pipeline 100 B2R_X=0
run ballot kitti_pipeline 100 B2R_LIB=$PWD/hdl_graph_slam_b200/_lib/alt/libb200reg_ballot.so
run base voxelgrid 200 B2R_X=0
run ballot voxelgrid 200 B2R_LIB=$PWD/hdl_graph_slam_b200/_lib/alt/libb200reg_ballot.so
timeout 600 python bench.py --steps 200 --warmup 5 --cpu-sample 0 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 env B2R_LIB=$PWD/hdl_graph_slam_b200/_lib/alt/libb200reg_ballot.so python bench.py --steps 200 --warmup 5 --cpu-sample 0 > $O/bench_n1_ballot.json 2> $O/bench_n1_ballot.err
for f in bench_n1 bench_n1_ballot; do python - <<PY
import json
d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["config"].get("strict_chain_value"), d["kernel_ms_in_timed_region"], (d.get("loop_batch_n1") or {}).get("value"))
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_odo.csv python bench.py --steps 4 --warmup 3 --no-profile --cpu-sample 0 --no-anchor > $O/ncu_odo.log 2>&1
timeout 300 env B2R_LIB=$PWD/hdl_graph_slam_b200/_lib/alt/libb200reg_ballot.so ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_odo_ballot.csv python bench.py --steps 4 --warmup 3 --no-profile --cpu-sample 0 --no-anchor > $O/ncu_odo_ballot.log 2>&1
