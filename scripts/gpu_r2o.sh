# Scaling points N = 2 and 4 of the bench contract (identical start states on every rank), for the 1/2/4/8 table.
mkdir -p gpurun_out
for n in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 2 --warmup 3 --no-c3 --no-c4 > gpurun_out/r2o_bench_${n}gpu.json 2> gpurun_out/r2o_bench_${n}gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2o_bench_${n}gpu.json').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ['value','n_gpus','ms_per_step','dimer_searches_per_s']}, d['e2e']['value'])"
done
