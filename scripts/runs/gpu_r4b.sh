#!/bin/bash
# round 2, run 4b: ncu --set full of the pair kernels and the row-order kernel on index-ordered rows
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_eam_rho|k_eam_force|k_nb_order_rows' -s 7 -c 3 -o gpurun_out/r4b_eam_pair python scripts/time_pairs.py > gpurun_out/r4b_ncu_eam.log 2>&1
ls -la gpurun_out/r4b*
