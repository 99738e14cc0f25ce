"""Profiling driver (run under ncu on the GPU box): one small dimer batch through the C ABI.
C3-style 6^3 Fe-H cell (431 atoms), R replicas. Usage: prof_search_small.py [R] [n]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfly_b200 import fly, inputs

R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
wen = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "wen.eam.fs")
ctx = fly.Context(0)
data = fly.DataEAM(ctx, wen)
pot = fly.Generic(data)
pos, ex, ids = inputs.iron_hydrogen_cell(n, 2.855, remove=(1,), hydrogen=True)
s1 = fly.System(data, fly.Box(tuple(ex)), ids)
rmin, _, failed, _ = fly.LBFGS().minimise(s1, pos, pot)
P, A = inputs.perturbed_batch(rmin[0], ex, len(ids) - 1, R, seed=5)
sR = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
for _ in range(2):
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(sR, P, A, rmin[0], pot)
print("exits", np.unique(code, return_counts=True), "evals", cnt[:, 0].sum())
