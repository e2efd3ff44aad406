#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -q -x --durations=12 -k "voxelgrid or prefilter or batch_gpu or abi" > $O/pytest_new.txt 2>&1; echo "pytest exit $?" >> $O/pytest_new.txt
tail -30 $O/pytest_new.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ndt_derivatives" -s 3 -c 2 -o $O/prof_ndt python tools/prof_ndt.py > $O/ncu_ndt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bvh_build_cluster|k_knn_cov_reg" -s 2 -c 2 -o $O/prof_build python tools/prof_one.py > $O/ncu_build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_search|k_pair_accumulate|k_knn_cov_reg_batch|k_bvh_build" -s 2 -c 6 -o $O/prof_batch python tools/prof_batch.py 8 1 > $O/ncu_batch.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_voxelgrid_cluster" -c 1 -o $O/prof_vg python -c "
import sys; sys.path.insert(0,'.')
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import synth
r=pkg.select_registration_method({'registration_method':'FAST_GICP'})
c=synth.scan('kitti',frame=1)
for _ in range(3): o=r.voxelGridFilter(c,0.25)
print(o.shape)
" > $O/ncu_vg.log 2>&1
