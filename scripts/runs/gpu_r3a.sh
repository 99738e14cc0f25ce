#!/bin/bash
# round 2, run a: full GPU test-suite, smoke, short bench (both register variants of pass 2)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r3a_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r3a_smoke.log 2>&1
python bench.py --steps 2 --warmup 1 > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
OFLY_FORCE_MINB=3 python bench.py --steps 1 --warmup 1 --no-cpu --no-c3 --no-c4 --no-c5 --no-c1 > gpurun_out/r3a_bench_minb3.json 2> gpurun_out/r3a_bench_minb3.err
tail -5 gpurun_out/r3a_tests.log; cat gpurun_out/r3a_smoke.log | tail -3; cut -c1-600 gpurun_out/r3a_bench.json
