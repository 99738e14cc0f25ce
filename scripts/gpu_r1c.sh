set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1c_pytest.log
tail -5 gpurun_out/r1c_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1c_bench.log 2>&1
tail -c 2500 gpurun_out/r1c_bench.log
OFLY_SEARCH_LOOP=host timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r1c_launches_bench.log 2>&1
tail -c 600 gpurun_out/r1c_launches_bench.log
