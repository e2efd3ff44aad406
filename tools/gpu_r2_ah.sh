#!/bin/bash
# 2 GPUs, final build: sharded batch == single GPU (bitwise) and the loop-batch bench line under torchrun
O=gpurun_out/r2ah; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > $O/check_multi.txt 2>&1; echo "check exit $?" >> $O/check_multi.txt; tail -2 $O/check_multi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_loop_n2.json 2> $O/bench_loop_n2.err
python - <<PY
import json
d=json.loads(open("$O/bench_loop_n2.json").read().strip().splitlines()[-1]); print("loop n2", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], (d.get("anchor_n1") or d.get("loop_batch_n1") or {}).get("value"), {k:v for k,v in (d.get("clocks") or {}).items() if k!="window"})
PY
