mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -c 400 gpurun_out/r2j_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ['value','ms_per_step','ticks_per_step']}, d['e2e']['value']); c=d.get('c3_environment'); print(c['c5_shard']); print({k:v for k,v in c.items() if k!='c5_shard'})"
