"""Numerical-sensitivity study (CPU only): does expanding the L-BFGS two-loop's dot products over Gram entries
(the single-pass formulation of search.cuh) change search outcomes? Builds the oracle twice (plain two-loop and
-DFLY_ORACLE_LBFGS_GRAM) and compares exit codes, work counters, saddle positions and barriers on the test cells."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.common import WEN, bcc_vacancy  # noqa: E402

out = os.path.join(tempfile.gettempdir(), "liboracle_gram.so")
subprocess.run(["/usr/bin/g++", "-std=c++17", "-O3", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-DFLY_ORACLE_LBFGS_GRAM",
                "-o", out, os.path.join(ROOT, "oracle", "capi.cpp")], check=True)

cells, reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 8
base, ex, ids = bcc_vacancy(cells, a=2.855, remove=(1,), hydrogen=True)
fz = np.zeros(len(ids), np.uint8)
res = {}
for name, path in (("two-loop", None), ("gram", out)):
    tab = O.Tables(WEN, libpath=path) if path else O.Tables(WEN)
    osys = O.System(tab, ex)
    r0, g0, failed, cnt = osys.minimise(base, ids, fz, threads=8)
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(reps)]
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    o_r, o_ax, code, c = osys.find_sp_batch(P, A, ids, fz, r0, threads=8)
    osys.rebuild(tab.rcut, r0)
    E0 = osys.energy(ids, 4)
    bar = []
    for k in range(reps):
        osys.rebuild(tab.rcut, o_r[k])
        bar.append(osys.energy(ids, 4) - E0)
    res[name] = (r0, cnt, code, c, o_r, np.array(bar))
    print(name, "minimise counters", cnt, "codes", code.tolist(), "evals", c[:, 0].tolist())
a, b = res["two-loop"], res["gram"]
print("min pos diff", np.abs(a[0] - b[0]).max())
print("codes equal", a[2].tolist() == b[2].tolist(), "evals total", a[3][:, 0].sum(), b[3][:, 0].sum())
ok = a[2] == 0
print("saddle pos diff (success)", [float(np.abs(a[4][k] - b[4][k]).max()) for k in range(reps) if ok[k] and b[2][k] == 0])
print("barrier diff (success)", [float(abs(a[5][k] - b[5][k])) for k in range(reps) if ok[k] and b[2][k] == 0])
