"""Timing probe (GPU box): wall time of one dimer batch. Usage: time_search.py R n [repeat]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfly_b200 import fly, inputs

R = int(sys.argv[1]); n = int(sys.argv[2]); rep = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wen = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "wen.eam.fs")
ctx = fly.Context(0)
data = fly.DataEAM(ctx, wen)
pot = fly.Generic(data)
pos, ex, ids = inputs.iron_hydrogen_cell(n, 2.855, remove=(1,), hydrogen=True)
s1 = fly.System(data, fly.Box(tuple(ex)), ids)
t0 = time.perf_counter(); rmin, _, failed, c1 = fly.LBFGS().minimise(s1, pos, pot); t1 = time.perf_counter() - t0
print(f"[{os.environ.get('OFLY_LBFGS', 'twoloop')} if={os.environ.get('OFLY_REBUILD_IF', '1')}] minimise N={len(ids)}: {t1*1e3:.1f} ms, evals {c1[0,0]}, {t1/c1[0,0]*1e6:.1f} us/tick")
P, A = inputs.perturbed_batch(rmin[0], ex, len(ids) - 1, R, seed=5)
sR = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
for _ in range(rep):
    l0 = ctx.tick_count()
    t0 = time.perf_counter()
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(sR, P, A, rmin[0], pot)
    dt = time.perf_counter() - t0
    ticks = ctx.tick_count() - l0
    ev = cnt[:, 0].sum()
    print(f"dimer R={R} N={len(ids)}: {dt*1e3:.1f} ms, ticks {ticks:.0f} ({dt/ticks*1e6:.1f} us/tick), evals {ev} max {cnt[:,0].max()}, "
          f"{ev*len(ids)/dt:.3e} atom-evals/s, exits {dict(zip(*np.unique(code, return_counts=True)))}")
if os.environ.get("OFLY_TRACE_LONGEST"):
    os.environ["OFLY_TRACE"] = str(int(np.argmax(cnt[:, 0])))
    print("tracing replica", os.environ["OFLY_TRACE"], "evals", cnt[:, 0].max(), flush=True)
    fly.Dimer().find_sp(sR, P, A, rmin[0], pot)
