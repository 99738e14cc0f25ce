# Lone 16000-atom replica: tick time by sub-warp width of the pair kernels.
for g in 0 8 16 32; do echo "== OFLY_GROUP=$g"; OFLY_GROUP=$g OFLY_TRACE_LONGEST=1 timeout 300 python scripts/time_search.py 1 20 2 2>&1 | grep -E "^dimer|91 -> 92|93 -> 94|90 -> 95|95 -> 91|94 ->  1"; done
