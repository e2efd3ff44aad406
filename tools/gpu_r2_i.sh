#!/bin/bash
# 2-GPU box: full GPU suite (multi-GPU tests included), multi-GPU bitwise check, loop-batch scaling N=1,2 and the reference arm
O=gpurun_out/r2i; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -15 $O/pytest_gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > $O/check_multi.txt 2>&1; echo "check exit $?" >> $O/check_multi.txt
tail -5 $O/check_multi.txt
timeout 400 python bench.py --workload loop_batch > $O/bench_loop_n1.json 2> $O/bench_loop_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $O/bench_loop_n2.json 2> $O/bench_loop_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
tail -2 $O/bench_loop_n1.json $O/bench_loop_n2.json $O/bench_ref_n2.json
