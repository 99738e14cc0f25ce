#!/bin/bash
# round 2, run k: cost of the literal two-loop L-BFGS against the Gram step on the C2 batch and the C5-sized batch
set -x
mkdir -p gpurun_out
(python scripts/time_search.py 64 20 2) > gpurun_out/r3k_c2_gram.log 2>&1
(OFLY_LBFGS=twoloop python scripts/time_search.py 64 20 2) > gpurun_out/r3k_c2_twoloop.log 2>&1
(python scripts/time_search.py 1024 6 2) > gpurun_out/r3k_c5_gram.log 2>&1
(OFLY_LBFGS=twoloop python scripts/time_search.py 1024 6 2) > gpurun_out/r3k_c5_twoloop.log 2>&1
tail -2 gpurun_out/r3k_c2_gram.log gpurun_out/r3k_c2_twoloop.log gpurun_out/r3k_c5_gram.log gpurun_out/r3k_c5_twoloop.log
