# C3-size (432 atoms, 256 replicas) tick-latency study: team width x sub-warp width x L-BFGS variant, and a launch list.
mkdir -p gpurun_out
for lb in gram twoloop; do for th in 128 256 512; do
  echo "== OFLY_LBFGS=$lb OFLY_TEAM_THREADS=$th"; OFLY_LBFGS=$lb OFLY_TEAM_THREADS=$th timeout 200 python scripts/time_search.py 256 6 1 2>&1 | tail -1
done; done
for g in 8 16 32; do echo "== OFLY_GROUP=$g threads 512"; OFLY_GROUP=$g OFLY_TEAM_THREADS=512 timeout 200 python scripts/time_search.py 256 6 1 2>&1 | tail -1; done
echo "== R=2048 threads 512 gram"; OFLY_TEAM_THREADS=512 timeout 300 python scripts/time_search.py 2048 6 1 2>&1 | tail -1
echo "== R=2048 threads 512 twoloop"; OFLY_LBFGS=twoloop OFLY_TEAM_THREADS=512 timeout 300 python scripts/time_search.py 2048 6 1 2>&1 | tail -1
echo "== R=2048 threads 128 twoloop"; OFLY_LBFGS=twoloop timeout 300 python scripts/time_search.py 2048 6 1 2>&1 | tail -1
OFLY_SEARCH_LOOP=host timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r1i_launches_c3.csv python scripts/prof_search_small.py 256 6 > gpurun_out/r1i_ncu_c3.log 2>&1
tail -c 300 gpurun_out/r1i_ncu_c3.log
