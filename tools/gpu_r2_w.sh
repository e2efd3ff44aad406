#!/bin/bash
# 1 GPU: warp-autonomous NDT pass (32-point sub-chunks, warp tickets, three-level tree) — suite, NDT bench A/B against the block-chunk build, ncu
O=gpurun_out/r2w; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so hdl_graph_slam_b200/_lib/alt/*.so > $O/lib.md5
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt
timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
timeout 600 env B2R_LIB=$PWD/hdl_graph_slam_b200/_lib/alt/libb200reg_blockchunks.so python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1_blockchunks.json 2> $O/bench_ndt_n1_blockchunks.err
for f in bench_ndt_n1 bench_ndt_n1_blockchunks; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("roofline") or {}).get("avg_launch_us"), d["config"].get("strict_chain_value"), d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ndt_derivatives" -s 3 -c 1 -o $O/prof_ndt python tools/prof_ndt.py > $O/ncu_full_ndt.log 2>&1
