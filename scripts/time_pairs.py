"""Tuning aid: time the EAM gradient launch pair and the neighbour rebuild on the C2 batch (64 x 16 000) and on C4 (one
247 500-atom system) with CUDA events. Run under different OFLY_LIB / OFLY_GROUP / OFLY_PAIR_CTAS settings to compare variants."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfly_b200 import capi, fly, inputs  # noqa: E402

WEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "wen.eam.fs")
ctx = fly.Context(0)
lib = ctx.lib
data = fly.DataEAM(ctx, WEN)
stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", 0))
r_list = 5.302 + (1.1 ** (1 / 3) - 1) * 5.302


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    ctx.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n):
        fn()
    b.record(stream)
    ctx.synchronize()
    return a.elapsed_time(b) / n


out = {"lib": os.environ.get("OFLY_LIB", "default"), "group": os.environ.get("OFLY_GROUP"), "minb": os.environ.get("OFLY_FORCE_MINB")}
base, ex, ids = inputs.iron_hydrogen_cell(20, 2.8553, remove=(1,), hydrogen=True)
R = 64
P = np.stack([inputs.perturb(base, ex, len(ids) - 1, np.random.default_rng([77, r]))[0] for r in range(R)])
s = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
d, g = ctx.alloc(P.nbytes), ctx.alloc(P.nbytes)
ctx.upload(d, P)
capi.check(lib.ofly_neigh_rebuild_dev(s.h, r_list, d))
out["c2_pair_ms"] = timed(lambda: capi.check(lib.ofly_eam_gradient_dev(s.h, g)), 30)
out["c2_rebuild_ms"] = timed(lambda: capi.check(lib.ofly_neigh_rebuild_dev(s.h, r_list, d)), 10, warm=2)
G = np.zeros_like(P)
ctx.download(G, g)
out["c2_grad_checksum"] = float(np.abs(G).sum())
ctx.free(d), ctx.free(g)
s.close()
p4, ex4, id4 = inputs.vacancy_cell(50, 2.8553, 0.01, 0.05, seed=4331)
s4 = fly.System(data, fly.Box(tuple(ex4)), id4)
d4, g4 = ctx.alloc(p4.nbytes), ctx.alloc(p4.nbytes)
ctx.upload(d4, p4)
capi.check(lib.ofly_neigh_rebuild_dev(s4.h, r_list, d4))
out["c4_pair_ms"] = timed(lambda: capi.check(lib.ofly_eam_gradient_dev(s4.h, g4)), 30)
out["c4_rebuild_ms"] = timed(lambda: capi.check(lib.ofly_neigh_rebuild_dev(s4.h, r_list, d4)), 10, warm=2)
print(json.dumps(out), flush=True)
