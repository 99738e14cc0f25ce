# Final round-1 evidence with the committed engine: GPU tests, smoke, bench (own arm + reference arm), C3/C5-sized timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1w_bench.json 2> gpurun_out/r1w_bench.err
tail -c 600 gpurun_out/r1w_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1w_bench_reference.json 2>> gpurun_out/r1w_bench.err
timeout 300 python scripts/time_search.py 256 6 2 2>&1 | tail -2
timeout 300 python scripts/time_search.py 2048 6 2 2>&1 | tail -2
for f in gpurun_out/r1w_bench.json gpurun_out/r1w_bench_reference.json; do python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ['impl','value','ms_per_step','ticks_per_step','mean_active_replicas','dimer_searches_per_s']}, d['e2e']['value'], d.get('roofline',{}).get('frac'), d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))"; done
