#!/bin/bash
# round 2, run 4n: loads in flight per thread and vector pass of the segment kernel (4 / 6 / 8 elements): lone 16 000-atom replica, bulk ticks
set -x
mkdir -p gpurun_out
: > gpurun_out/r4n_pass_unroll.txt
for v in default pu6 pu8; do
  if [ $v = default ]; then unset OFLY_LIB; else export OFLY_LIB=$PWD/openfly_b200/variants/libopenfly_b200_$v.so; fi
  timeout 120 python scripts/time_search.py 1 20 3 2>&1 | grep dimer | tail -2 | sed "s/^/[$v] /" >> gpurun_out/r4n_pass_unroll.txt
  timeout 120 python scripts/time_bulk.py 40 3 2>&1 | tail -1 | sed "s/^/[$v] /" >> gpurun_out/r4n_pass_unroll.txt
done
cat gpurun_out/r4n_pass_unroll.txt
