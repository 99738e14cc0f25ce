#!/bin/bash
# round 2, run q: 8 GPUs -- the bench as the driver launches it
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r3q_bench_8gpu.json 2> gpurun_out/r3q_bench_8gpu.err
tail -c 600 gpurun_out/r3q_bench_8gpu.json; tail -5 gpurun_out/r3q_bench_8gpu.err
