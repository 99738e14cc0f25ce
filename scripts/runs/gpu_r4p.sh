#!/bin/bash
# round 2, run 4p: final state -- the whole -m gpu suite, smoke, the default bench line, ncu of the rebuild kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4p_tests.log 2>&1
tail -3 gpurun_out/r4p_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4p_smoke.txt 2>&1; tail -1 gpurun_out/r4p_smoke.txt
timeout 900 python bench.py > gpurun_out/r4p_bench.json 2> gpurun_out/r4p_bench.err
cut -c1-300 gpurun_out/r4p_bench.json
ncu --set full --clock-control none --import-source on -k regex:'k_nb_build|k_nb_order_rows|k_nb_sort_cells' -s 3 -c 3 -o gpurun_out/r4p_rebuild python scripts/time_pairs.py > gpurun_out/r4p_ncu_rebuild.log 2>&1
