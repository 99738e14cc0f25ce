#!/bin/bash
# round 2, run 4c: register bitonic row sort (rebuild timing), segment CTAs of 256 threads (two per SM) against 512
set -x
mkdir -p gpurun_out
python scripts/time_pairs.py > gpurun_out/r4c_pairs.json 2> gpurun_out/r4c_pairs.err
cat gpurun_out/r4c_pairs.json
python scripts/time_search.py 64 20 2 > gpurun_out/r4c_search_512.txt 2>&1
OFLY_SEG_THREADS=256 python scripts/time_search.py 64 20 2 > gpurun_out/r4c_search_256.txt 2>&1
OFLY_SEG_THREADS=256 OFLY_CLUSTER=8 python scripts/time_search.py 64 20 2 > gpurun_out/r4c_search_256c8.txt 2>&1
tail -3 gpurun_out/r4c_search_*.txt
