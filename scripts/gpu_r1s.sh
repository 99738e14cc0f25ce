# Scalar trig on one warp per CTA: parity + C3 timings/trace.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
OFLY_TRACE_LONGEST=1 timeout 300 python scripts/time_search.py 256 6 1 2>&1 | grep -E "dimer|90 ->|2 -> 10|20 -> 21|23 -> 24|12 -> 13|10 -> 11|22 -> 23"
timeout 300 python scripts/time_search.py 2048 6 2 2>&1 | tail -2
