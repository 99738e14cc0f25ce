#!/bin/bash
# round 2, run i: full GPU suite after the f3/f4 entry points
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r3i_tests.log
tail -14 gpurun_out/r3i_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
