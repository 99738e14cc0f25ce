set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r1a_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eam -s 4 -c 2 -o gpurun_out/r1a_prof_grad python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r1a_prof_grad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_search -s 2 -c 1 -o gpurun_out/r1a_prof_search python scripts/prof_search_small.py 128 6 > gpurun_out/r1a_prof_search.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nb_ -s 12 -c 6 -o gpurun_out/r1a_prof_nb python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r1a_prof_nb.log 2>&1
ls -la gpurun_out
