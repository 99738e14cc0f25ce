#!/bin/bash
# round 2, run u: full GPU suite (every pytest file under its own timeout), smoke
set -x
mkdir -p gpurun_out
for f in tests/test_gpu_eam_neigh.py tests/test_gpu_search.py tests/test_gpu_mech.py tests/test_gpu_full_size.py tests/test_cpp_shim.py tests/test_golden_fixtures.py; do
  timeout 300 python -m pytest $f -m gpu -x -q 2>&1 | tail -4
done
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
