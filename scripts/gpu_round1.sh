set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/r1_bench.log 2>&1
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r1_ref.log 2>&1
tail -3 gpurun_out/r1_pytest.log gpurun_out/r1_smoke.log gpurun_out/r1_bench.log gpurun_out/r1_ref.log
