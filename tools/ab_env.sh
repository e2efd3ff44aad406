#!/bin/bash
gicp() { python bench.py --steps ${STEPS:-100} --warmup 5 --cpu-sample 0 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('gicp', round(d['value'],1), round(d['e2e']['value'],1), d['kernel_ms_in_timed_region'], d['config']['mean_iterations'])"; }
L=hdl_graph_slam_b200/_lib
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python tools/knn_time.py
for lib in main c2 c8; do
  [ $lib != main ] && cp $L/alt/libb200reg_$lib.so $L/libb200reg.so
  echo "== $lib"; python tools/odo_times.py 12 q; gicp
done
cp $L/alt/libb200reg_knnprof.so $L/libb200reg.so
python tools/prof_odo.py 3 | tail -1
