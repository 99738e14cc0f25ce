#!/bin/bash
# round 2, run m: profiles -- launch list of a bench step (host-launched ticks), ncu --set full of the pair kernels (bench geometry),
# of a full tick (segment + rebuild kernels), and of the batched eigenvalue kernel
set -x
mkdir -p gpurun_out
OFLY_SEARCH_LOOP=host ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 6000 --csv --log-file gpurun_out/r3m_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-c1 --no-c3 --no-c4 --no-c5 > gpurun_out/r3m_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_eam_rho|k_eam_force' -s 6 -c 2 -o gpurun_out/r3m_eam_pair python scripts/time_pairs.py > gpurun_out/r3m_ncu_eam.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_nb_build|k_nb_sort_cells|k_nb_canon_count|k_nb_scatter' -s 4 -c 4 -o gpurun_out/r3m_rebuild python scripts/time_pairs.py > gpurun_out/r3m_ncu_rebuild.log 2>&1
OFLY_SEARCH_LOOP=host ncu --set full --clock-control none --import-source on -k regex:'k_segment' -s 6 -c 2 -o gpurun_out/r3m_segment python scripts/time_search.py 64 20 1 > gpurun_out/r3m_ncu_segment.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_sym_eigenvalues|k_hess_columns' -c 3 -o gpurun_out/r3m_eigen python -m pytest tests/test_gpu_mech.py -m gpu -x -q -k every_saddle > gpurun_out/r3m_ncu_eigen.log 2>&1
ls -la gpurun_out/r3m*
