#!/bin/bash
set -x
mkdir -p gpurun_out
OFLY_VERBOSE=1 python bench.py --steps 1 --warmup 1 --no-cpu --no-c1 --no-c5 --no-c4 > gpurun_out/r3o_a.json 2> gpurun_out/r3o_a.err
OFLY_VERBOSE=1 python bench.py --steps 1 --warmup 1 --no-cpu --no-c1 --no-c5 > gpurun_out/r3o_b.json 2> gpurun_out/r3o_b.err
OFLY_VERBOSE=1 python bench.py --steps 1 --warmup 1 --no-cpu --no-c1 --no-c4 > gpurun_out/r3o_c.json 2> gpurun_out/r3o_c.err
for f in a b c; do python - <<PY
import json
d=json.loads(open('gpurun_out/r3o_$f.json').read().strip().splitlines()[-1])
c=d['c3_environment']; print('$f', c.get('hessian_eigen_s'), c.get('saddles'), c.get('search_s'))
PY
grep -c "own kernel" gpurun_out/r3o_$f.err; grep "eigenvalues of" gpurun_out/r3o_$f.err | sort | uniq -c | head -5
done
