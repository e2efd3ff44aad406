#!/bin/bash
# 1 GPU: the whole GPU suite on the final tree
O=gpurun_out/r2aj; mkdir -p $O
timeout 120 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
