#!/bin/bash
# round 2, run 4g: fused cooperative rebuild kernel against the IF node (tail and bulk ticks), then the whole -m gpu suite
set -x
mkdir -p gpurun_out
for f in 0 1; do OFLY_REBUILD_FUSED=$f timeout 120 python scripts/time_search.py 1024 6 1 2>&1 | grep dimer | sed "s/^/[fused $f] /"; done > gpurun_out/r4g_tail.txt
for f in 0 1; do OFLY_REBUILD_FUSED=$f timeout 120 python scripts/time_search.py 1 20 1 2>&1 | grep dimer | sed "s/^/[fused $f] /"; done >> gpurun_out/r4g_tail.txt
for f in 0 1; do OFLY_REBUILD_FUSED=$f timeout 120 python scripts/time_bulk.py 40 3 2>&1 | tail -1 | sed "s/^/[fused $f] /"; done >> gpurun_out/r4g_tail.txt
cat gpurun_out/r4g_tail.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4g_tests.log 2>&1
tail -8 gpurun_out/r4g_tests.log
