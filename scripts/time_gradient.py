"""Timing probe (GPU box): EAM gradient (k_eam_rho + k_eam_force) on a resident batch. Usage: time_gradient.py R n [iters]
OFLY_GROUP pins the sub-warp width."""
import ctypes as C, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openfly_b200 import capi, fly, inputs

R = int(sys.argv[1]); n = int(sys.argv[2]); iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
wen = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "wen.eam.fs")
ctx = fly.Context(0); data = fly.DataEAM(ctx, wen)
pos, ex, ids = inputs.iron_hydrogen_cell(n, 2.8553, remove=(1,), hydrogen=True)
rng = np.random.default_rng(0)
P = np.stack([pos + rng.normal(0, 0.02, pos.shape) for _ in range(R)])
s = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
d_pos, d_out = ctx.alloc(P.nbytes), ctx.alloc(P.nbytes)
ctx.upload(d_pos, P)
capi.check(ctx.lib.ofly_neigh_rebuild_dev(s.h, 5.3 + 0.002 + 0.1711, d_pos))
stream = torch.cuda.ExternalStream(ctx.stream())
for _ in range(3): capi.check(ctx.lib.ofly_eam_gradient_dev(s.h, d_out))
ctx.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(iters): capi.check(ctx.lib.ofly_eam_gradient_dev(s.h, d_out))
e1.record(stream); ctx.synchronize()
ms = e0.elapsed_time(e1) / iters
nbar = fly.List(s, 0).mean_count()
print(f"gradient R={R} N={len(ids)} G={os.environ.get('OFLY_GROUP','auto')}: {ms*1e3:.1f} us/eval-batch, {R*len(ids)/ms*1e-6:.4f} G atom-evals/s, nbar {nbar:.2f}, "
      f"roofline frac {R*len(ids)/(ms*1e-3)*(8*nbar+64)/6555.2e9:.4f}")
