# Two-GPU run of the bench contract (torchrun, one rank per GPU, NCCL only for the barrier / timing reduction).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r1r_bench_2gpu.json 2> gpurun_out/r1r_bench_2gpu.err
tail -c 1500 gpurun_out/r1r_bench_2gpu.json; tail -c 500 gpurun_out/r1r_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r1r_bench_ref_2gpu.json 2>> gpurun_out/r1r_bench_2gpu.err
tail -c 600 gpurun_out/r1r_bench_ref_2gpu.json
