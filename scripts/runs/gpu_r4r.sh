#!/bin/bash
# round 2, run 4r: last check of the clean rebuild -- smoke and the static-path + search tests
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4r_smoke.txt 2>&1; tail -1 gpurun_out/r4r_smoke.txt
timeout 600 python -m pytest tests/test_gpu_eam_neigh.py tests/test_gpu_search.py -m gpu -x -q > gpurun_out/r4r_tests.log 2>&1
tail -2 gpurun_out/r4r_tests.log
