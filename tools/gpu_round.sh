#!/bin/bash
# one gpurun call: parity tests, bench (both arms), ncu launch list.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "=== bench b200"; python bench.py --steps ${STEPS:-100} --warmup 5 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_err.log
echo "=== bench reference"; python bench.py --impl reference --steps 10 --warmup 2 2>&1 | tee gpurun_out/bench_ref.json | cut -c1-400
if [ -n "$NCU_LIST" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --cpu-sample 0 --no-profile > gpurun_out/ncu_bench.log 2>&1
  tail -2 gpurun_out/ncu_bench.log | cut -c1-300
fi
