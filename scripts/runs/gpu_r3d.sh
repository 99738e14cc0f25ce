#!/bin/bash
# round 2, run d: tile path -- tests, ncu of the tile kernels vs row kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r3d_tests.log
tail -15 gpurun_out/r3d_tests.log
ncu --set full --clock-control none --import-source on -k regex:'k_tile_eam|k_tile_build' -s 6 -c 3 -o gpurun_out/r3d_tile python scripts/time_pairs.py > gpurun_out/r3d_ncu_tile.log 2>&1
OFLY_TILES=0 ncu --set full --clock-control none --import-source on -k regex:'k_eam_rho|k_eam_force' -s 4 -c 2 -o gpurun_out/r3d_row python scripts/time_pairs.py > gpurun_out/r3d_ncu_row.log 2>&1
ls -la gpurun_out/*.ncu-rep
