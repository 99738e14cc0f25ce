#!/bin/bash
# round 2, run e: search step with the tile path and with the row kernels
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --no-c4 --no-c1 > gpurun_out/r3e_bench_tiles.json 2> gpurun_out/r3e_bench_tiles.err
OFLY_TILES=0 python bench.py --steps 2 --warmup 1 --no-cpu --no-c4 --no-c1 > gpurun_out/r3e_bench_rows.json 2> gpurun_out/r3e_bench_rows.err
OFLY_TILES=0 OFLY_FORCE_MINB=3 python scripts/time_pairs.py 2>&1 | tail -1 > gpurun_out/r3e_variants.jsonl
cut -c1-300 gpurun_out/r3e_bench_tiles.json; cut -c1-300 gpurun_out/r3e_bench_rows.json
