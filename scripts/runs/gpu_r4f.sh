#!/bin/bash
# round 2, run 4f: the whole -m gpu suite on index-ordered rows + 4-tick WHILE body, then a short bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4f_tests.log 2>&1
tail -8 gpurun_out/r4f_tests.log
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-c1 --no-c3 --no-c4 > gpurun_out/r4f_bench.json 2> gpurun_out/r4f_bench.err
tail -c 1500 gpurun_out/r4f_bench.json
