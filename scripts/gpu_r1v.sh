# Where does the time between two segments go? Stamps at the start/end of block 0 of the pair kernels; single replica.
OFLY_TRACE_LONGEST=1 timeout 300 python scripts/time_search.py 1 6 2 2>&1 | grep -E "dimer|9. ->|-> 9.|->  1"
OFLY_REBUILD_IF=0 OFLY_TRACE_LONGEST=1 timeout 300 python scripts/time_search.py 1 6 2 2>&1 | grep -E "dimer|9. ->|-> 9.|->  1"
