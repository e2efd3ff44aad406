#!/bin/bash
# 1 GPU: ncu --set full of the k-NN covariance kernels of the final build (single cloud and batched), and of the single-pair LM kernel
O=gpurun_out/r2af; mkdir -p $O
md5sum hdl_graph_slam_b200/_lib/libb200reg.so > $O/lib.md5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_knn_cov_reg|k_pair_lm" -s 2 -c 4 -o $O/prof_one_knn python tools/prof_one.py > $O/ncu_one.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_knn_cov_reg_batch" -c 1 -o $O/prof_batch_knn python tools/prof_batch.py 8 1 > $O/ncu_batch.log 2>&1
ls -la $O
