#!/bin/bash
# round 2, run 4o: branch-free candidate loop of the neighbour build (type bits carried by the image records): timing + set parity
set -x
mkdir -p gpurun_out
python scripts/time_pairs.py > gpurun_out/r4o_pairs.json 2> gpurun_out/r4o_pairs.err
cat gpurun_out/r4o_pairs.json
timeout 600 python -m pytest tests/test_gpu_eam_neigh.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r4o_tests.log 2>&1
tail -4 gpurun_out/r4o_tests.log
