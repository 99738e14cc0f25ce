#!/bin/bash
# round 2, run 4a: rows ordered by neighbour index in the transposed round layout -- pair / rebuild timing + parity of the static paths
set -x
mkdir -p gpurun_out
python scripts/time_pairs.py > gpurun_out/r4a_pairs.json 2> gpurun_out/r4a_pairs.err
cat gpurun_out/r4a_pairs.json
timeout 600 python -m pytest tests/test_gpu_eam_neigh.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r4a_tests.log 2>&1
tail -15 gpurun_out/r4a_tests.log
