#!/bin/bash
# 1 GPU: full suite on the new build; A/B of batch lanes and search occupancy on the loop batch; NDT persistent grid; new workloads
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -8 $O/pytest_gpu.txt
run() { name=$1; shift; timeout 400 env "$@" python bench.py --workload loop_batch --no-profile > $O/loop_$name.json 2> $O/loop_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/loop_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"])
except Exception as e: print("$name ERR", e)
PY
}
run lanes2 B2R_BATCH_LANES=2
run lanes1 B2R_BATCH_LANES=1
run lanes2_mb10 B2R_BATCH_LANES=2 B2R_LIB=$PWD/hdl_graph_slam_b200/_lib/alt/libb200reg_mb10.so
run lanes2_mb12 B2R_BATCH_LANES=2 B2R_LIB=$PWD/hdl_graph_slam_b200/_lib/alt/libb200reg_mb12.so
run lanes2_c2 B2R_BATCH_LANES=2 B2R_BATCH_COPIES=2
timeout 400 python bench.py --workload loop_batch > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
timeout 600 env B2R_NDT_WAVES=100 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 > $O/bench_ndt_n1_oldgrid.json 2> $O/bench_ndt_n1_oldgrid.err
timeout 600 env B2R_NDT_WAVES=0.5 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 > $O/bench_ndt_n1_half.json 2> $O/bench_ndt_n1_half.err
timeout 600 python bench.py --workload voxelgrid --steps 200 --warmup 5 > $O/bench_voxelgrid.json 2> $O/bench_voxelgrid.err
timeout 900 python bench.py --workload kitti_pipeline --steps 100 --warmup 5 > $O/bench_kitti.json 2> $O/bench_kitti.err
timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
for f in bench_ndt_n1 bench_ndt_n1_oldgrid bench_ndt_n1_half bench_voxelgrid bench_kitti bench_n1; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("roofline") or {}).get("avg_launch_us"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
