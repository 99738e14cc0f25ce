# Where does a tail tick go? Graph-loop overhead probe, C3 launch list (host loop), C3 timings by team width.
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/graph_loop_probe scripts/graph_loop_probe.cu && /tmp/graph_loop_probe
for lb in gram twoloop; do for th in 128 512; do
  echo "== OFLY_LBFGS=$lb OFLY_TEAM_THREADS=$th"; OFLY_LBFGS=$lb OFLY_TEAM_THREADS=$th timeout 200 python scripts/time_search.py 256 6 1 2>&1 | tail -1
done; done
echo "== host loop, gram 512"; OFLY_SEARCH_LOOP=host timeout 200 python scripts/time_search.py 256 6 1 2>&1 | tail -1
echo "== flat graph (no IF), gram 512"; OFLY_REBUILD_IF=0 timeout 200 python scripts/time_search.py 256 6 1 2>&1 | tail -1
OFLY_SEARCH_LOOP=host timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r1k_launches_c3.csv python scripts/prof_search_small.py 256 6 > gpurun_out/r1k_ncu_c3.log 2>&1
tail -c 200 gpurun_out/r1k_ncu_c3.log
