#!/bin/bash
# shortest possible A/B of the experimental oriented-box variant: standalone correspondence-kernel time per odometry frame, then parity
L=hdl_graph_slam_b200/_lib
echo "== main"; python tools/odo_times.py 10 q
cp $L/alt/libb200reg_obb.so $L/libb200reg.so
echo "== obb"; python tools/odo_times.py 10 q
python -m pytest tests/test_gicp_gpu.py tests/test_callers_gpu.py -m gpu -q -x 2>&1 | tail -2
