# Neighbour-build sub-warp width in the search ticks (the rows do not depend on it): bench ms/step by OFLY_BUILD_WAVES.
for w in 1 2 4 8 16; do
OFLY_BUILD_WAVES=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-c3 --no-c4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('waves=$w', d['ms_per_step'], d['ticks_per_step'], d['value'])"
done
