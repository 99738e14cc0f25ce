#!/bin/bash
# round 2, run j: Gram vs two-loop branch-outs against the oracle (256 C3 searches)
set -x
mkdir -p gpurun_out
python scripts/lbfgs_branch_study.py 256 2>&1 | tail -3 > gpurun_out/r3j_lbfgs.jsonl
cat gpurun_out/r3j_lbfgs.jsonl
