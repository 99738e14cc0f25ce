# ncu --set full of one late tick of a small batch (8 replicas x 431 atoms): the latency regime of the tick kernels.
mkdir -p gpurun_out
OFLY_SEARCH_LOOP=host timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_(segment|eam_rho_dyn|eam_force_dyn)$' -s 900 -c 3 -o gpurun_out/r2l_tail_tick python scripts/prof_search_small.py 8 6 > gpurun_out/r2l_ncu.log 2>&1
tail -c 300 gpurun_out/r2l_ncu.log
