#!/bin/bash
# round 2, run b: GPU tests (mech API, libfly surface), pair-kernel variants
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r3b_tests.log
for v in default row0 row2 knotint shiftb all3 plain; do
  if [ $v = default ]; then unset OFLY_LIB; else export OFLY_LIB=$PWD/openfly_b200/variants/libopenfly_b200_$v.so; fi
  python scripts/time_pairs.py 2>&1 | tail -1 >> gpurun_out/r3b_variants.jsonl
done
export OFLY_LIB=$PWD/openfly_b200/variants/libopenfly_b200_all3.so
for g in 8 16; do OFLY_GROUP=$g python scripts/time_pairs.py 2>&1 | tail -1 >> gpurun_out/r3b_variants.jsonl; done
OFLY_FORCE_MINB=3 python scripts/time_pairs.py 2>&1 | tail -1 >> gpurun_out/r3b_variants.jsonl
OFLY_PAIR_CTAS=16 python scripts/time_pairs.py 2>&1 | tail -1 >> gpurun_out/r3b_variants.jsonl
unset OFLY_LIB
tail -12 gpurun_out/r3b_tests.log; cat gpurun_out/r3b_variants.jsonl
