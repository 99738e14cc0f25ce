#!/bin/bash
# round 2, run 4m: conditional handles set only on change -- late ticks, then the search tests (device loop against host loop, queue against batch)
set -x
mkdir -p gpurun_out
timeout 120 python scripts/time_search.py 1024 6 1 2>&1 | grep dimer > gpurun_out/r4m_tail.txt
timeout 120 python scripts/time_search.py 1 20 2 2>&1 | grep dimer >> gpurun_out/r4m_tail.txt
timeout 120 python scripts/time_bulk.py 40 3 2>&1 | tail -1 >> gpurun_out/r4m_tail.txt
cat gpurun_out/r4m_tail.txt
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_full_size.py tests/test_gpu_mech.py -m gpu -x -q > gpurun_out/r4m_tests.log 2>&1
tail -4 gpurun_out/r4m_tests.log
