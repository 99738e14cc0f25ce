#!/bin/bash
# round 2, run f: own eigenvalue kernel -- mech tests, C3 leg timing
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mech.py tests/test_gpu_eam_neigh.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r3f_tests.log
tail -12 gpurun_out/r3f_tests.log
python bench.py --steps 1 --warmup 1 --no-cpu --no-c4 --no-c1 --no-c5 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3f_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('c3_environment')))
PY
tail -3 gpurun_out/r3f_bench.err
