#!/bin/bash
# round 2, run w: why is the C3 prefactor leg 4x slower inside the full default bench? (verbose phase timings)
set -x
mkdir -p gpurun_out
OFLY_VERBOSE=1 timeout 900 python bench.py > gpurun_out/r3w_bench.json 2> gpurun_out/r3w_bench.err
grep "openfly_b200:" gpurun_out/r3w_bench.err | grep -v "search graph" | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3w_bench.json').read().strip().splitlines()[-1])
print(d['c3_environment'])
PY
