#!/bin/bash
# One-call validation on a B200 box (gpurun -- 'bash tools/gpu_validate.sh [N]'): what the driver runs at round end, plus every bench workload.
#   N = 1 (default): pytest -m gpu, smoke(), bench.py (both arms), the other workloads, ncu launch lists of the three main workloads
#   N > 1          : pytest -m gpu (multi-GPU tests included), tools/check_multi_gpu.py, bench.py --gpus N (both arms)
N=${1:-1}
O=gpurun_out/validate_n$N; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so > $O/lib.md5
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -4 $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
show() { python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d.get('impl', 'b200'), 'value', round(d.get('value', float('nan')), 1), 'e2e', round((d.get('e2e') or {}).get('value', float('nan')), 1), 'ms/step', d.get('ms_per_step'),
              'roofline', (d.get('roofline') or {}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'anchor', (d.get('anchor_n1') or d.get('loop_batch_n1') or {}).get('value'),
              'strict', (d.get('config') or {}).get('strict_chain_value'), d.get('unavailable', ''))
    except Exception as e:
        print(f, 'ERR', e)
PY
}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err
  timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
  timeout 600 python bench.py --workload loop_batch > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
  timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
  timeout 600 python bench.py --workload voxelgrid --steps 200 --warmup 5 > $O/bench_voxelgrid.json 2> $O/bench_voxelgrid.err
  timeout 900 python bench.py --workload kitti_pipeline --steps 100 --warmup 5 > $O/bench_kitti.json 2> $O/bench_kitti.err
  show $O/bench_ref_n1.json $O/bench_n1.json $O/bench_loop_n1.json $O/bench_ndt_n1.json $O/bench_voxelgrid.json $O/bench_kitti.json
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_odo.csv python bench.py --steps 4 --warmup 3 --no-profile --cpu-sample 0 --no-anchor > $O/ncu_odo.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_batch.csv python bench.py --workload loop_batch --steps 1 --warmup 1 --pairs 64 --no-profile --cpu-sample 0 > $O/ncu_batch.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_ndt.csv python bench.py --workload ndt_odometry_hdl32e_128k --steps 4 --warmup 3 --no-profile --cpu-sample 0 --no-anchor > $O/ncu_ndt.log 2>&1
else
  P=29500
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) tools/check_multi_gpu.py > $O/check_multi.txt 2>&1; echo "check exit $?" >> $O/check_multi.txt
  tail -2 $O/check_multi.txt
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+3)) bench.py --gpus $N > $O/bench_loop_n$N.json 2> $O/bench_loop_n$N.err
  show $O/bench_ref_n$N.json $O/bench_loop_n$N.json
fi
