#!/bin/bash
# round 2, run 4e: WHILE body of 4 ticks against 1 (late ticks of small and large replicas); rebuild without the in-cell sort;
# team shapes for a lone 16 000-atom replica
set -x
mkdir -p gpurun_out
python scripts/time_pairs.py > gpurun_out/r4e_pairs.json 2> gpurun_out/r4e_pairs.err
cat gpurun_out/r4e_pairs.json
for u in 1 4 8; do OFLY_TICK_UNROLL=$u python scripts/time_search.py 1024 6 1 2>&1 | grep dimer | sed "s/^/[unroll $u] /"; done > gpurun_out/r4e_tail.txt
for u in 1 4; do OFLY_TICK_UNROLL=$u python scripts/time_search.py 1 20 1 2>&1 | grep dimer | sed "s/^/[unroll $u] /"; done >> gpurun_out/r4e_tail.txt
OFLY_CLUSTER=8 python scripts/time_search.py 1 20 1 2>&1 | grep dimer | sed "s/^/[cluster 8] /" >> gpurun_out/r4e_tail.txt
OFLY_SEG_THREADS=256 python scripts/time_search.py 1 20 1 2>&1 | grep dimer | sed "s/^/[threads 256] /" >> gpurun_out/r4e_tail.txt
cat gpurun_out/r4e_tail.txt
