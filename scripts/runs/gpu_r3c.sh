#!/bin/bash
# round 2, run c: tile path -- tests, timings tiles vs rows, short bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r3c_tests.log
tail -15 gpurun_out/r3c_tests.log
python scripts/time_pairs.py 2>&1 | tail -1 > gpurun_out/r3c_variants.jsonl
OFLY_TILES=0 python scripts/time_pairs.py 2>&1 | tail -1 >> gpurun_out/r3c_variants.jsonl
OFLY_TILE_CTAS=6 python scripts/time_pairs.py 2>&1 | tail -1 >> gpurun_out/r3c_variants.jsonl
OFLY_TILE_CTAS=24 python scripts/time_pairs.py 2>&1 | tail -1 >> gpurun_out/r3c_variants.jsonl
cat gpurun_out/r3c_variants.jsonl
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
cut -c1-400 gpurun_out/r3c_bench.json; tail -3 gpurun_out/r3c_bench.err
