#!/bin/bash
gicp() { python bench.py --steps ${STEPS:-100} --warmup 5 --cpu-sample 0 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('gicp', round(d['value'],1), round(d['e2e']['value'],1), d['kernel_ms_in_timed_region'], d['config']['mean_iterations'])"; }
ndt() { python bench.py --workload ndt_odometry_hdl32e_128k --steps 30 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ndt', round(d['value'],1), round(d['e2e']['value'],1), d['kernel_ms_in_timed_region'])"; }
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
echo "== all on"; gicp; gicp
echo "== no index seed"; B2R_NO_INDEX_SEED=1 gicp
echo "== no pdl"; B2R_NO_PDL=1 gicp
ndt
python bench.py --workload loop_batch --pairs 128 2>/dev/null | cut -c1-120
