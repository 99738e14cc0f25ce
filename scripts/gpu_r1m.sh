# Warp-parallel recursion, single-trip pass B, torque/dg fusion, one team per list entry, dynamic sub-warp width.
mkdir -p gpurun_out
OFLY_VERBOSE=1 timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
OFLY_TRACE_LONGEST=1 timeout 300 python scripts/time_search.py 256 6 1 2>&1 | tail -32
OFLY_LBFGS=twoloop timeout 300 python scripts/time_search.py 256 6 1 2>&1 | tail -1
timeout 300 python scripts/time_search.py 2048 6 1 2>&1 | tail -1
OFLY_TRACE_LONGEST=1 timeout 600 python scripts/time_search.py 16 20 1 2>&1 | tail -32
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err
tail -c 1700 gpurun_out/r1m_bench.json; tail -c 600 gpurun_out/r1m_bench.err
