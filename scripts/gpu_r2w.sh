# Default bench of both arms on the committed tree (final line of the round).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
tail -c 300 gpurun_out/r2w_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/r2w_bench_reference.json 2>> gpurun_out/r2w_bench.err
for f in gpurun_out/r2w_bench.json gpurun_out/r2w_bench_reference.json; do python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ['impl','value','ms_per_step','ticks_per_step','dimer_searches_per_s','gpu_launches']}, d['e2e']['value'], d.get('roofline',{}).get('frac'), d.get('cpu_baseline'))"; done
