# In-kernel time stamps of the longest replica: where a tick goes (C3 and C2 sizes).
mkdir -p gpurun_out
OFLY_TRACE_LONGEST=1 timeout 300 python scripts/time_search.py 256 6 1 2>&1 | tail -40
OFLY_TRACE_LONGEST=1 OFLY_LBFGS=twoloop timeout 300 python scripts/time_search.py 256 6 1 2>&1 | tail -30
OFLY_TRACE_LONGEST=1 timeout 600 python scripts/time_search.py 16 20 1 2>&1 | tail -40
