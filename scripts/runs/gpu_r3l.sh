#!/bin/bash
# round 2, run l: staged device loop -- tests, C2 / C5 timings staged vs single graph
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3l_tests.log
tail -5 gpurun_out/r3l_tests.log
(python scripts/time_search.py 64 20 2) 2>&1 | tail -n 1
(OFLY_STAGES=0 python scripts/time_search.py 64 20 2) 2>&1 | tail -n 1
(python scripts/time_search.py 1024 6 2) 2>&1 | tail -n 1
(OFLY_STAGES=0 python scripts/time_search.py 1024 6 2) 2>&1 | tail -n 1
(python scripts/time_search.py 256 6 2) 2>&1 | tail -n 1
(OFLY_STAGES=0 python scripts/time_search.py 256 6 2) 2>&1 | tail -n 1
