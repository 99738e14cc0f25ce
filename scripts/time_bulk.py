"""Tuning aid (GPU box): cost of a search tick with every slot busy. 64 dimer searches on the 16 000-atom cell are cut at
a small max_steps, so nearly every slot is active in every tick; prints time per tick and per replica-evaluation.
Usage: time_bulk.py [max_steps=20] [repeat=3]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfly_b200 import fly, inputs

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wen = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "wen.eam.fs")
ctx = fly.Context(0)
data = fly.DataEAM(ctx, wen)
pot = fly.Generic(data)
pos, ex, ids = inputs.iron_hydrogen_cell(20, 2.855, remove=(1,), hydrogen=True)
R = 64
P, A = inputs.perturbed_batch(pos, ex, len(ids) - 1, R, seed=5)
sR = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
dim = fly.Dimer(fly.dimer_options(max_steps=steps, convex_max=1000))
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("OFLY_"))
for _ in range(rep):
    l0 = ctx.tick_count()
    t0 = time.perf_counter()
    out_r, out_ax, code, cnt = dim.find_sp(sR, P, A, pos, pot)
    dt = time.perf_counter() - t0
    ticks = ctx.tick_count() - l0
    ev = int(cnt[:, 0].sum())
    print(f"[{tag}] bulk R={R} N={len(ids)}: {dt*1e3:.1f} ms, ticks {ticks:.0f} ({dt/ticks*1e6:.1f} us/tick), evals {ev} ({ev/ticks:.1f} per tick, "
          f"{dt/ev*1e6:.2f} us per replica-eval), {ev*len(ids)/dt:.3e} atom-evals/s", flush=True)
