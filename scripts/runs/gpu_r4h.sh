#!/bin/bash
# round 2, run 4h: pass 1 capped for 5 / 6 CTAs per SM against uncapped (60 registers, 4 CTAs); default bench line
set -x
mkdir -p gpurun_out
python scripts/time_pairs.py > gpurun_out/r4h_pairs.jsonl 2> gpurun_out/r4h_pairs.err
for v in rho5 rho6; do OFLY_LIB=$PWD/openfly_b200/variants/libopenfly_b200_$v.so python scripts/time_pairs.py >> gpurun_out/r4h_pairs.jsonl 2>> gpurun_out/r4h_pairs.err; done
cat gpurun_out/r4h_pairs.jsonl
timeout 900 python bench.py > gpurun_out/r4h_bench.json 2> gpurun_out/r4h_bench.err
tail -c 3000 gpurun_out/r4h_bench.json
