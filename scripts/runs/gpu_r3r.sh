#!/bin/bash
# round 2, run r: entry prefetch + register-resident two-loop -- tests, pair timing, step timing with / without
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_search.py tests/test_golden_fixtures.py -m gpu -x -q 2>&1 | tail -6
python scripts/time_pairs.py 2>&1 | tail -n 1
(python scripts/time_search.py 64 20 2) 2>&1 | tail -n 1
(OFLY_LBFGS_REG=0 python scripts/time_search.py 64 20 2) 2>&1 | tail -n 1
(python scripts/time_search.py 1024 6 2) 2>&1 | tail -n 1
(OFLY_LBFGS_REG=0 python scripts/time_search.py 1024 6 2) 2>&1 | tail -n 1
