#!/bin/bash
# round 2, run v: the default bench line and the reference arm (queue of 512), as the driver runs them
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3v_bench_reference.json 2> gpurun_out/r3v_bench_reference.err
timeout 900 python bench.py > gpurun_out/r3v_bench.json 2> gpurun_out/r3v_bench.err
cut -c1-200 gpurun_out/r3v_bench_reference.json; cut -c1-200 gpurun_out/r3v_bench.json; tail -3 gpurun_out/r3v_bench.err
