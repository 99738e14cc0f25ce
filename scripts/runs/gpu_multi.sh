#!/bin/bash
# round 2, run 4k (was y): the sharded bench on N GPUs as the driver launches it (N from the environment)
set -x
mkdir -p gpurun_out
N=${OFLY_NGPU:-8}
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 2 --warmup 1 --no-c1 --no-c4 > gpurun_out/r4k_bench_${N}gpu.json 2> gpurun_out/r4k_bench_${N}gpu.err
tail -c 300 gpurun_out/r4k_bench_${N}gpu.json; tail -3 gpurun_out/r4k_bench_${N}gpu.err
