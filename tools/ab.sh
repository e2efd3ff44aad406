#!/bin/bash
run() { echo "== $1"; shift; env "$@" python bench.py --workload ndt_odometry_hdl32e_128k --steps 30 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ndt', round(d['value'],1), d['kernel_ms_in_timed_region'], d['clocks'])"; env "$@" python bench.py --workload loop_batch --pairs 64 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('loop', round(d['value'],1), d['mean_iterations'])"; }
run default A=1
run nopriority B2R_NO_PRIORITY=1
run spin_unbounded B2R_SPIN_LIMIT=4000000000
run spin_none B2R_SPIN_LIMIT=1
