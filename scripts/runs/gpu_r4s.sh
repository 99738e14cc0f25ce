#!/bin/bash
# round 2, run 4s: ncu launch list of a bench step with the final engine (ticks launched from the host: same kernels, same bits)
set -x
mkdir -p gpurun_out
OFLY_SEARCH_LOOP=host timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 6000 --csv --log-file gpurun_out/r4s_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-c1 --no-c3 --no-c4 --no-c5 > gpurun_out/r4s_ncu_launches.log 2>&1
wc -l gpurun_out/r4s_launches_bench.csv
