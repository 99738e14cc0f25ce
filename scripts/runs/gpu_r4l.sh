#!/bin/bash
# round 2, run 4l: in-kernel time stamps of late ticks with the current engine (literal two-loop, 4-tick body)
set -x
mkdir -p gpurun_out
OFLY_TRACE_LONGEST=1 timeout 200 python scripts/time_search.py 1024 6 1 > gpurun_out/r4l_anatomy_1024_c3.txt 2>&1
OFLY_TRACE_LONGEST=1 timeout 200 python scripts/time_search.py 1 20 1 > gpurun_out/r4l_anatomy_lone_c2.txt 2>&1
tail -40 gpurun_out/r4l_anatomy_1024_c3.txt
