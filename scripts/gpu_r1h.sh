# Round-1 (session 6): single-pass L-BFGS step + IF-guarded rebuild. Parity tests, A/B timings, bench, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
OFLY_VERBOSE=1 timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/time_search.py 256 6
OFLY_REBUILD_IF=0 timeout 300 python scripts/time_search.py 256 6 1
OFLY_LBFGS=twoloop timeout 300 python scripts/time_search.py 256 6 1
timeout 300 python scripts/time_search.py 2048 6 1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err
tail -c 3500 gpurun_out/r1h_bench.json; tail -c 600 gpurun_out/r1h_bench.err
OFLY_LBFGS=twoloop timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1h_bench_twoloop.json 2>> gpurun_out/r1h_bench.err
tail -c 3500 gpurun_out/r1h_bench_twoloop.json
OFLY_SEARCH_LOOP=host timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r1h_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r1h_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r1h_ncu_bench.log
