#!/bin/bash
# 1 GPU: the detect() plan / batch / replay test (host logic added after the validated build; device code unchanged: SASS identical)
O=gpurun_out/r2ai; mkdir -p $O
timeout 150 python -m pytest tests/test_batch_gpu.py -m gpu -q -x -k "detect_plan or equals_each_pair" > $O/pytest.txt 2>&1; echo "pytest exit $?" >> $O/pytest.txt; tail -3 $O/pytest.txt
