#!/bin/bash
# round 2, run p: full GPU suite after the f3 export / ownership changes
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r3p_tests.log
tail -14 gpurun_out/r3p_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
