# Round-1 re-measurement of the lock-step engine: GPU parity tests, smoke, timings, bench lines, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 300 python scripts/time_search.py 256 6
timeout 300 python scripts/time_search.py 2048 6 1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r1g_bench.json 2> gpurun_out/r1g_bench.err
tail -c 3000 gpurun_out/r1g_bench.json; tail -c 600 gpurun_out/r1g_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r1g_bench_reference.json 2>> gpurun_out/r1g_bench.err
tail -c 1500 gpurun_out/r1g_bench_reference.json
OFLY_SEARCH_LOOP=host timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r1g_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r1g_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r1g_ncu_bench.log
