#!/bin/bash
# 1 GPU: NDT pass with an even point split over the persistent blocks — NDT tests, bench, A/B against the chunked split
O=gpurun_out/r2s; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so > $O/lib.md5
timeout 600 python -m pytest tests -m gpu -q -x -k "ndt or parity_sizes or golden" > $O/pytest_ndt.txt 2>&1; echo "pytest exit $?" >> $O/pytest_ndt.txt
tail -3 $O/pytest_ndt.txt
timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
timeout 600 env B2R_NDT_WAVES=0.75 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1_w075.json 2> $O/bench_ndt_n1_w075.err
timeout 600 env B2R_NDT_WAVES=0.5 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1_w05.json 2> $O/bench_ndt_n1_w05.err
for f in bench_ndt_n1 bench_ndt_n1_w075 bench_ndt_n1_w05; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("roofline") or {}).get("avg_launch_us"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ndt_derivatives" -s 3 -c 1 -o $O/prof_ndt python tools/prof_ndt.py > $O/ncu_full_ndt.log 2>&1
