timeout 120 python scripts/time_search.py 1 20 || echo "R=1 FAILED/TIMEOUT"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/time_search.py 64 20
timeout 300 python scripts/time_search.py 256 6
OFLY_CLUSTER=1 timeout 300 python scripts/time_search.py 1 20
OFLY_CLUSTER=2 timeout 300 python scripts/time_search.py 1 20
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1f_bench.log 2>&1
tail -c 1800 gpurun_out/r1f_bench.log
