# Grid size of the pair kernels (CTAs per SM of the grid-stride launch): rounding-neutral tuning on the bench.
for w in 32 128 1024; do
OFLY_PAIR_CTAS=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-c3 --no-c4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pair ctas/SM=$w', round(d['ms_per_step'],1), d['ticks_per_step'], round(d['roofline']['launch_ms'],4), round(d['roofline']['frac'],4))"
done
