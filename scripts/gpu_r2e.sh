# 8-GPU sanity run of the bench contract (short): one rank per GPU, weak scaling.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 2 --warmup 3 --no-c3 --no-c4 > gpurun_out/r2e_bench_8gpu.json 2> gpurun_out/r2e_bench_8gpu.err
tail -c 300 gpurun_out/r2e_bench_8gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_8gpu.json').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ['value','n_gpus','ms_per_step','dimer_searches_per_s','scaling']}, d['e2e']['value'])"
