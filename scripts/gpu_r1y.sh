# One-wave rule for the dynamic sub-warp width: parity + timings.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/time_search.py 1 20 2 2>&1 | tail -1
timeout 300 python scripts/time_search.py 256 6 2 2>&1 | tail -1
timeout 300 python scripts/time_search.py 2048 6 2 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r1y_bench.json 2> gpurun_out/r1y_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r1y_bench.json').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ['value','ms_per_step','ticks_per_step','mean_active_replicas','dimer_searches_per_s']}, d['e2e']['value'], d['roofline']['frac'])"
