# Cluster-16 teams for large systems: parity tests, C2 trace and bench with 16 vs 8 CTAs per team.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
OFLY_TRACE_LONGEST=1 timeout 600 python scripts/time_search.py 16 20 1 2>&1 | grep -E "dimer|90 ->|2 -> 10|20 -> 21|23 -> 24|12 -> 13"
OFLY_CLUSTER16=0 OFLY_TRACE_LONGEST=1 timeout 600 python scripts/time_search.py 16 20 1 2>&1 | grep -E "dimer|90 ->|2 -> 10|20 -> 21|23 -> 24|12 -> 13"
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1p_bench.json 2> gpurun_out/r1p_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r1p_bench.json').read().strip().splitlines()[-1]); print('c16', d['value'], d['ms_per_step'], d['ticks_per_step'], d['mean_active_replicas'], d['e2e']['value'])"
OFLY_CLUSTER16=0 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1p_bench_c8.json 2>> gpurun_out/r1p_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r1p_bench_c8.json').read().strip().splitlines()[-1]); print('c8 ', d['value'], d['ms_per_step'], d['ticks_per_step'], d['mean_active_replicas'], d['e2e']['value'])"
tail -c 400 gpurun_out/r1p_bench.err
