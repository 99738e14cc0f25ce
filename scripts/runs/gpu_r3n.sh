#!/bin/bash
# round 2, run n: the default bench line and the reference arm, as the driver runs them
set -x
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3n_bench_reference.json 2> gpurun_out/r3n_bench_reference.err
python bench.py > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err
cut -c1-300 gpurun_out/r3n_bench_reference.json; cut -c1-300 gpurun_out/r3n_bench.json; tail -3 gpurun_out/r3n_bench.err
