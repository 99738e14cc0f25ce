#!/bin/bash
# round 2, run 4i: smoke, pair / rebuild timing, ncu --set full of the pair kernels and the rebuild kernels (bench geometry), default bench line
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4i_smoke.txt 2>&1; tail -1 gpurun_out/r4i_smoke.txt
python scripts/time_pairs.py > gpurun_out/r4i_pairs.json 2> gpurun_out/r4i_pairs.err; cat gpurun_out/r4i_pairs.json
ncu --set full --clock-control none --import-source on -k regex:'k_eam_rho|k_eam_force' -s 6 -c 2 -o gpurun_out/r4i_eam_pair python scripts/time_pairs.py > gpurun_out/r4i_ncu_eam.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_nb_build|k_nb_order_rows|k_nb_sort_cells' -s 3 -c 3 -o gpurun_out/r4i_rebuild python scripts/time_pairs.py > gpurun_out/r4i_ncu_rebuild.log 2>&1
timeout 900 python bench.py > gpurun_out/r4i_bench.json 2> gpurun_out/r4i_bench.err
cut -c1-600 gpurun_out/r4i_bench.json
ls -la gpurun_out/r4i*
