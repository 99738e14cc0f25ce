"""Which L-BFGS arithmetic should be the default? The 64-search C3 set of tests/test_gpu_search.py (and 192 more start states)
run with the literal two-loop recursion (default) and with the two-pass Gram step (OFLY_LBFGS=gram), each classified
against the CPU oracle: same exit and same saddle, or branched out. Prints one JSON line per mode with counts and wall time."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openfly_b200 import fly  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.common import WEN, bcc_vacancy  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
O.build()
otab = O.Tables(WEN)
base, ex, ids = bcc_vacancy(6, a=2.855, remove=(1,), hydrogen=True)
fz = np.zeros(len(ids), np.uint8)
osys = O.System(otab, ex)
r0, _, failed, _ = osys.minimise(base, ids, fz, threads=8)
starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(R)]
P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
t0 = time.perf_counter()
o_r, o_ax, o_code, o_cnt = osys.find_sp_batch(P, A, ids, fz, r0, threads=os.cpu_count())
t_cpu = time.perf_counter() - t0
ctx = fly.Context(0)
data = fly.DataEAM(ctx, WEN)
for mode in ("gram", "twoloop"):
    if mode == "gram":
        os.environ["OFLY_LBFGS"] = "gram"
    else:
        os.environ.pop("OFLY_LBFGS", None)
    s = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
    fly.Dimer().find_sp(s, P, A, r0, fly.Generic(data))
    t0 = time.perf_counter()
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(s, P, A, r0, fly.Generic(data))
    dt = time.perf_counter() - t0
    s.close()
    exits_differ = int((code != o_code).sum())
    other_saddle = int(sum(1 for k in range(R) if code[k] == o_code[k] == 0 and np.abs(out_r[k] - o_r[k]).max() > 1e-4))
    short = [int(min(cnt[k, 0], o_cnt[k, 0])) for k in range(R)
             if code[k] != o_code[k] or (code[k] == 0 and np.abs(out_r[k] - o_r[k]).max() > 1e-4)]
    print(json.dumps({"mode": mode, "searches": R, "branch_outs": exits_differ + other_saddle, "exit_differs": exits_differ,
                      "other_saddle": other_saddle, "shortest_branched_trajectory_evals": (min(short) if short else None),
                      "gpu_seconds": dt, "gpu_evals": int(cnt[:, 0].sum()), "oracle_evals": int(o_cnt[:, 0].sum()),
                      "cpu_seconds_all_cores": t_cpu}), flush=True)
