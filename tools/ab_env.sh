#!/bin/bash
gicp() { python bench.py --steps ${STEPS:-100} --warmup 5 --cpu-sample 0 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('gicp', round(d['value'],1), round(d['e2e']['value'],1), d['kernel_ms_in_timed_region'], d['config']['mean_iterations'])"; }
loop() { python bench.py --workload loop_batch --pairs 256 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('loop', round(d['value'],1), d['ms_per_step'], d['mean_iterations'])"; }
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
gicp
echo "== loop default"; loop; loop
echo "== loop no pdl"; B2R_NO_PDL=1 loop
echo "== loop no index seed"; B2R_NO_INDEX_SEED=1 loop
