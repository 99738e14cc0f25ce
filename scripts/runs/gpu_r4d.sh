#!/bin/bash
# round 2, run 4d: tick cost with every slot busy -- segment CTAs of 512 threads (one per SM) against 256 (two per SM)
set -x
mkdir -p gpurun_out
python scripts/time_bulk.py 40 3 > gpurun_out/r4d_bulk.txt 2>&1
OFLY_SEG_THREADS=256 python scripts/time_bulk.py 40 3 >> gpurun_out/r4d_bulk.txt 2>&1
OFLY_SEG_THREADS=256 OFLY_CLUSTER=8 python scripts/time_bulk.py 40 3 >> gpurun_out/r4d_bulk.txt 2>&1
OFLY_CLUSTER=8 python scripts/time_bulk.py 40 3 >> gpurun_out/r4d_bulk.txt 2>&1
OFLY_LBFGS=gram python scripts/time_bulk.py 40 3 >> gpurun_out/r4d_bulk.txt 2>&1
cat gpurun_out/r4d_bulk.txt
