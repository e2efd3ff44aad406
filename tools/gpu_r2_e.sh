#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
B2R_DEBUG_BUILD=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python tools/prof_batch.py 1 1 > $O/sanitizer_batch.log 2>&1
head -40 $O/sanitizer_batch.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python tools/prof_one.py > $O/sanitizer_single.log 2>&1
tail -5 $O/sanitizer_single.log
