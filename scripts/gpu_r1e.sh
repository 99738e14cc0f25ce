timeout 300 python scripts/time_search.py 1 20
OFLY_SEARCH_LOOP=host timeout 300 python scripts/time_search.py 1 20
timeout 300 python scripts/time_search.py 256 6
OFLY_SEARCH_LOOP=host timeout 300 python scripts/time_search.py 256 6
OFLY_SEARCH_LOOP=host timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/r1e_launches_small.csv python scripts/time_search.py 256 6 1 > /dev/null 2>&1
OFLY_SEARCH_LOOP=host timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/r1e_launches_r1.csv python scripts/time_search.py 1 20 1 > /dev/null 2>&1
