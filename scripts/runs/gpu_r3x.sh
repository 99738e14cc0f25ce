#!/bin/bash
# round 2, run x: padded Hessian stride -- mech tests, then the C3 prefactor leg three times in full default benches
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mech.py tests/test_gpu_eam_neigh.py -m gpu -x -q 2>&1 | tail -3
for k in 1 2; do
OFLY_VERBOSE=1 timeout 900 python bench.py > gpurun_out/r3x_bench_$k.json 2> gpurun_out/r3x_bench_$k.err
grep "Hessians assembled" gpurun_out/r3x_bench_$k.err | tail -1
done
