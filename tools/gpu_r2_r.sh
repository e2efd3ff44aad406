#!/bin/bash
# 1 GPU: faster hilbert30, first search beside the source k-NN covariance kernel (strict chain), optional 16-CTA clusters — suite, benches, phase cycles
O=gpurun_out/r2r; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so > $O/lib.md5
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -4 $O/pytest_gpu.txt
timeout 400 python bench.py --workload loop_batch --cpu-sample 0 > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 600 python bench.py --steps 200 --warmup 5 --cpu-sample 0 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1.json 2> $O/bench_ndt_n1.err
timeout 600 python bench.py --workload voxelgrid --steps 200 --warmup 5 --cpu-sample 0 > $O/bench_voxelgrid.json 2> $O/bench_voxelgrid.err
timeout 900 python bench.py --workload kitti_pipeline --steps 100 --warmup 5 --cpu-sample 0 > $O/bench_kitti.json 2> $O/bench_kitti.err
for f in bench_loop_n1 bench_n1 bench_ndt_n1 bench_voxelgrid bench_kitti; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d["config"].get("strict_chain_value"), d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
export B2R_LIB=$PWD/hdl_graph_slam_b200/_lib/alt/libb200reg_buildprof.so
timeout 120 python tools/prof_one.py > $O/buildprof_64k.txt 2>&1
timeout 120 python - > $O/buildprof_misc.txt 2>&1 <<PY
import sys; sys.path.insert(0,'.')
import numpy as np
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import synth
r=pkg.select_registration_method({'registration_method':'FAST_GICP'})
c=synth.scan('kitti',frame=1)
v=r.voxelGridFilter(c,0.25)
for n in (v.shape[0], 16000, 131072):
    a=synth.scan('hdl32e',frame=2)[:n] if n>v.shape[0] else v
    r.setInputTarget(a); r.synchronize()
    print('built', a.shape)
PY
grep -h "^build" $O/buildprof_64k.txt | tail -3; grep -h "^build" $O/buildprof_misc.txt | tail -6
timeout 600 env B2R_NO_COV_OVERLAP=1 python bench.py --steps 200 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_n1_nooverlap.json 2> $O/bench_n1_nooverlap.err
timeout 600 env B2R_NO_CLUSTER16=1 python bench.py --workload voxelgrid --steps 200 --warmup 5 --cpu-sample 0 > $O/bench_voxelgrid_cl8.json 2> $O/bench_voxelgrid_cl8.err
timeout 600 env B2R_NO_CLUSTER16=1 python bench.py --workload ndt_odometry_hdl32e_128k --steps 50 --warmup 5 --cpu-sample 0 --no-anchor > $O/bench_ndt_n1_cl8.json 2> $O/bench_ndt_n1_cl8.err
for f in bench_n1_nooverlap bench_voxelgrid_cl8 bench_ndt_n1_cl8; do python - <<PY
import json
try:
    d=json.loads(open("$O/$f.json").read().strip().splitlines()[-1]); print("$f", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"], d["config"].get("strict_chain_value"), d.get("kernel_ms_in_timed_region"))
except Exception as e: print("$f ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
