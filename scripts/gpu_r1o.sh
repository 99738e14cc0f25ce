# Shell-ordered neighbour rows: parity tests, gradient kernel timing, rebuild timing, bench, ncu of the pair kernels.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for g in 4 8; do OFLY_GROUP=$g timeout 200 python scripts/time_gradient.py 64 20; done
timeout 300 python scripts/time_search.py 256 6 1 2>&1 | tail -1
timeout 300 python scripts/time_search.py 2048 6 1 2>&1 | tail -1
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1o_bench.json 2> gpurun_out/r1o_bench.err
tail -c 1700 gpurun_out/r1o_bench.json; tail -c 600 gpurun_out/r1o_bench.err
OFLY_SEARCH_LOOP=host timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_(eam_(rho|force)|nb_sort_rows|nb_build)$' -c 4 -o gpurun_out/r1o_eam_pair python scripts/time_gradient.py 64 20 2 > gpurun_out/r1o_ncu_eam.log 2>&1
tail -c 200 gpurun_out/r1o_ncu_eam.log
