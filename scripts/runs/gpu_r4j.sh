#!/bin/bash
# round 2, run 4j: history vectors prefetched into L1 ahead of the two-loop recursion's reductions -- late ticks, bulk ticks, search parity
set -x
mkdir -p gpurun_out
timeout 120 python scripts/time_search.py 1024 6 1 2>&1 | grep dimer > gpurun_out/r4j_tail.txt
timeout 120 python scripts/time_search.py 1 20 2 2>&1 | grep dimer >> gpurun_out/r4j_tail.txt
timeout 120 python scripts/time_bulk.py 40 3 2>&1 | tail -1 >> gpurun_out/r4j_tail.txt
cat gpurun_out/r4j_tail.txt
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r4j_tests.log 2>&1
tail -6 gpurun_out/r4j_tests.log
