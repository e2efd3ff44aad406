#!/bin/bash
# 1 GPU: the driver's own invocations on the final tree (clean rebuild): pytest -m gpu, smoke(), reference arm, bench at --steps 20 --warmup 5
O=gpurun_out/r2ag; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so > $O/lib.md5
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt; tail -2 $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json; echo
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; python - <<PY
import json
d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1]); print("bench", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["config"].get("strict_chain_value"), {k:v for k,v in d["clocks"].items() if k!="window"}, d["gpu_launches"])
PY
