#!/bin/bash
# round 2, run g: two GPUs -- sharded bench (both arms) as the driver launches it
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 1 --warmup 1 --no-c4 --no-c1 --no-c3 > gpurun_out/r3g_bench_2gpu.json 2> gpurun_out/r3g_bench_2gpu.err
tail -c 1500 gpurun_out/r3g_bench_2gpu.json; tail -5 gpurun_out/r3g_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r3g_ref_2gpu.json 2> gpurun_out/r3g_ref_2gpu.err
cut -c1-300 gpurun_out/r3g_ref_2gpu.json; tail -3 gpurun_out/r3g_ref_2gpu.err
