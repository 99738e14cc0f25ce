#!/bin/bash
# round 2, run h: lone-replica tick anatomy (trace marks) for C2- and C3-sized systems, capped vs uncapped pass 2
set -x
mkdir -p gpurun_out
(OFLY_TRACE_LONGEST=1 python scripts/time_search.py 1 20 1) > gpurun_out/r3h_lone_c2.log 2>&1
(OFLY_TRACE_LONGEST=1 python scripts/time_search.py 1 6 1) > gpurun_out/r3h_lone_c3.log 2>&1
(OFLY_LIB=$PWD/openfly_b200/variants/libopenfly_b200_dyn3.so python scripts/time_search.py 1 20 1) > gpurun_out/r3h_lone_c2_dyn3.log 2>&1
(OFLY_LIB=$PWD/openfly_b200/variants/libopenfly_b200_dyn3.so python scripts/time_search.py 64 20 1) > gpurun_out/r3h_64_c2_dyn3.log 2>&1
(python scripts/time_search.py 64 20 1) > gpurun_out/r3h_64_c2.log 2>&1
(OFLY_TRACE_LONGEST=1 python scripts/time_search.py 1024 6 1) > gpurun_out/r3h_1024_c3.log 2>&1
tail -40 gpurun_out/r3h_lone_c2.log; tail -40 gpurun_out/r3h_lone_c3.log; cat gpurun_out/r3h_lone_c2_dyn3.log gpurun_out/r3h_64_c2_dyn3.log gpurun_out/r3h_64_c2.log; tail -40 gpurun_out/r3h_1024_c3.log
