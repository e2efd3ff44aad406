#!/bin/bash
# A/B of engine variants on one box: the main build first (tests + benches), then each _lib/alt/libb200reg_<name>.so swapped in.
mkdir -p gpurun_out
L=hdl_graph_slam_b200/_lib
cp $L/libb200reg.so /tmp/main.so
gicp() { python bench.py --steps ${STEPS:-100} --warmup 5 --cpu-sample 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('gicp', round(d['value'],1), round(d['e2e']['value'],1), d['kernel_ms_in_timed_region'], d['config']['mean_iterations'])"; }
ndt() { python bench.py --workload ndt_odometry_hdl32e_128k --steps 30 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ndt', round(d['value'],1), d['kernel_ms_in_timed_region'])"; }
echo "== main"; python -m pytest tests -m gpu -q -x 2>&1 | tail -4
gicp; gicp; ndt
for v in "$@"; do
  echo "== $v"; cp $L/alt/libb200reg_$v.so $L/libb200reg.so
  case $v in ndt*) ndt;; *) gicp; gicp;; esac
done
cp /tmp/main.so $L/libb200reg.so
