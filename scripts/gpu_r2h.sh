# Round-1 evidence for the committed tree: launch list of the bench command (host-driven tick loop so that ncu sees
# every kernel), and ncu --set full captures of the dominant kernels.
mkdir -p gpurun_out
OFLY_SEARCH_LOOP=host timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2h_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-c3 --no-c4 > gpurun_out/r2h_ncu_launches.log 2>&1
tail -c 200 gpurun_out/r2h_ncu_launches.log
OFLY_SEARCH_LOOP=host timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_eam_(rho|force)$' -c 2 -o gpurun_out/r2h_eam_pair python bench.py --steps 1 --warmup 1 --no-cpu --no-c3 --no-c4 > gpurun_out/r2h_ncu_eam.log 2>&1
tail -c 200 gpurun_out/r2h_ncu_eam.log
OFLY_SEARCH_LOOP=host timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^k_(segment|eam_rho_dyn|eam_force_dyn|nb_build_dyn)$' -s 480 -c 4 -o gpurun_out/r2h_tick python bench.py --steps 1 --warmup 1 --no-cpu --no-c3 --no-c4 > gpurun_out/r2h_ncu_tick.log 2>&1
tail -c 200 gpurun_out/r2h_ncu_tick.log
ls -la gpurun_out | tail -8
