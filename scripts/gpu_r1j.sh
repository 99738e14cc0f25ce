# G-invariant summation + dynamic sub-warp width, segment kernel over the active list with wide teams, flag-based list compaction.
mkdir -p gpurun_out
OFLY_VERBOSE=1 timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for g in 4 8 16; do OFLY_GROUP=$g timeout 200 python scripts/time_gradient.py 64 20; done
timeout 300 python scripts/time_search.py 256 6 1
OFLY_LBFGS=twoloop timeout 300 python scripts/time_search.py 256 6 1
timeout 300 python scripts/time_search.py 2048 6 1
OFLY_LBFGS=twoloop timeout 300 python scripts/time_search.py 2048 6 1
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err
tail -c 1500 gpurun_out/r1j_bench.json; tail -c 600 gpurun_out/r1j_bench.err
OFLY_LBFGS=twoloop timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1j_bench_twoloop.json 2>> gpurun_out/r1j_bench.err
tail -c 1500 gpurun_out/r1j_bench_twoloop.json
