#!/bin/bash
# multi-GPU legs (launched the way the driver does): odometry replicas + loop-closure batch with the NCCL all-gather
N=${N:-2}
mkdir -p gpurun_out
echo "=== odometry replicas x$N"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-60} --warmup 5 2> gpurun_out/bench_n${N}_err.log | tee gpurun_out/bench_n${N}.json | cut -c1-400
tail -2 gpurun_out/bench_n${N}_err.log
echo "=== loop batch x$N"
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload loop_batch --pairs ${PAIRS:-64} 2> gpurun_out/bench_loop_n${N}_err.log | tee gpurun_out/bench_loop_n${N}.json | cut -c1-900
tail -2 gpurun_out/bench_loop_n${N}_err.log
