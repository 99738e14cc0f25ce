#!/bin/bash
# round 2, run 4q: programmatic dependent launch edges pass 1 -> pass 2 -> next segment: late ticks, bulk ticks, search parity
set -x
mkdir -p gpurun_out
: > gpurun_out/r4q_pdl.txt
for p in 0 1; do
  OFLY_PDL=$p OFLY_VERBOSE=1 timeout 120 python scripts/time_search.py 1024 6 1 2>&1 | grep "dimer\|openfly_b200:" | sort -u | sed "s/^/[pdl $p] /" >> gpurun_out/r4q_pdl.txt
  OFLY_PDL=$p timeout 120 python scripts/time_search.py 1 20 3 2>&1 | grep dimer | tail -2 | sed "s/^/[pdl $p] /" >> gpurun_out/r4q_pdl.txt
  OFLY_PDL=$p timeout 120 python scripts/time_bulk.py 40 3 2>&1 | tail -1 | sed "s/^/[pdl $p] /" >> gpurun_out/r4q_pdl.txt
done
cat gpurun_out/r4q_pdl.txt
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q > gpurun_out/r4q_tests.log 2>&1
tail -3 gpurun_out/r4q_tests.log
