set -x
timeout 60 ./build/cond_graph_probe
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r1b_pytest.log
tail -30 gpurun_out/r1b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1
tail -5 gpurun_out/r1b_smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r1b_bench.log 2>&1
tail -c 3000 gpurun_out/r1b_bench.log
