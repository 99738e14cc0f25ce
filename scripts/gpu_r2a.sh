mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 400 gpurun_out/r2a_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ['value','ms_per_step','ticks_per_step']}, d['e2e']['value']); print(d['neighbour_rebuild']); print(d.get('c4_single_system'))"
