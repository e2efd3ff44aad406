#!/bin/bash
# 2 GPUs: suite (multi-GPU tests included) on the build with up-front prefetches in the accumulate kernel; loop batch N = 1 / 2 (identical shards), N = 2 with distinct shards + host timing
O=gpurun_out/r2l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -4 $O/pytest_gpu.txt
timeout 400 python bench.py --workload loop_batch > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $O/bench_loop_n2.json 2> $O/bench_loop_n2.err
timeout 400 env B2R_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --distinct-shards --no-profile --steps 3 > $O/bench_loop_n2_distinct.json 2> $O/bench_loop_n2_distinct.err
timeout 600 python bench.py --steps 200 --warmup 5 --no-anchor > $O/bench_n1.json 2> $O/bench_n1.err
for f in bench_loop_n1 bench_loop_n2 bench_loop_n2_distinct bench_n1; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d.get("per_rank_ms_per_step"), d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
grep "rank" $O/bench_loop_n2_distinct.err | tail -24
