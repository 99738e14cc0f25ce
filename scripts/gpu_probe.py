"""Development probe (run under gpurun): search kernels vs the oracle + quick timings. Not a test."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from openfly_b200 import fly
from tests.common import WEN, bcc_vacancy

ctx = fly.Context(0)
data = fly.DataEAM(ctx, WEN)
otab = O.Tables(WEN)
pot = fly.Generic(data)

def probe(n, R, a=2.855):
    base, ex, ids = bcc_vacancy(n, a=a, remove=(1,), hydrogen=True)
    N = len(ids)
    fz = np.zeros(N, np.uint8)
    osys = O.System(otab, ex)
    # relax the base cell on the CPU oracle and on the GPU
    t0 = time.time(); r0, g0, fail0, c0 = osys.minimise(base, ids, fz, threads=8); t_cpu = time.time() - t0
    s1 = fly.System(data, fly.Box(tuple(ex)), ids, replicas=1)
    t0 = time.time(); r1, g1, fail1, c1 = fly.LBFGS().minimise(s1, base, pot); t_gpu = time.time() - t0
    print(f"[minimise n={n} N={N}] cpu {t_cpu:.3f}s fail={fail0} cnt={c0} | gpu {t_gpu:.3f}s fail={fail1} cnt={c1[0]}")
    print("   |r_gpu - r_cpu| max", np.abs(r1[0] - r0).max(), " |g| cpu", np.linalg.norm(g0), "gpu", np.linalg.norm(g1))
    osys.rebuild(otab.rcut, r0); E0 = osys.energy(ids, 8)
    nl = fly.List(s1, data.rcut()); nl.rebuild(r1); E1 = pot.energy(nl)[0]
    print("   E_min cpu", E0, "gpu", E1, "diff", E1 - E0)
    s1.close()
    # perturbed starts centred on the H atom
    centre = N - 1
    starts = [osys.perturb(r0, fz, centre, (1, 3, 3, 4), rep) for rep in range(R)]
    P = np.stack([s[0] for s in starts]); A = np.stack([s[1] for s in starts])
    sR = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
    dim = fly.Dimer()
    t0 = time.time(); gr, gax, gcode, gcnt = dim.find_sp(sR, P, A, r0, pot); t_gpu = time.time() - t0
    t0 = time.time(); gr, gax, gcode, gcnt = dim.find_sp(sR, P, A, r0, pot); t_gpu2 = time.time() - t0
    nref = min(R, 16)
    t0 = time.time(); cr, cax, ccode, ccnt = osys.find_sp_batch(P[:nref], A[:nref], ids, fz, r0, threads=8); t_cpu = time.time() - t0
    print(f"[dimer n={n} N={N} R={R}] gpu {t_gpu:.3f}s (2nd {t_gpu2:.3f}s) total grad evals {gcnt[:,0].sum()} | cpu {nref} searches {t_cpu:.3f}s evals {ccnt[:,0].sum()}")
    print("   gpu atom-evals/s", gcnt[:, 0].sum() * N / t_gpu2, " cpu atom-evals/s", ccnt[:, 0].sum() * N / t_cpu)
    print("   exit gpu", gcode[:nref].tolist()); print("   exit cpu", ccode.tolist())
    print("   evals gpu", gcnt[:nref, 0].tolist()); print("   evals cpu", ccnt[:, 0].tolist())
    for k in range(nref):
        if gcode[k] == 0 and ccode[k] == 0:
            d = np.abs(gr[k] - cr[k]).max()
            osys.rebuild(otab.rcut, cr[k]); Ec = osys.energy(ids, 8)
            osys.rebuild(otab.rcut, gr[k]); Eg = osys.energy(ids, 8)
            print(f"   rep {k}: |dr|max {d:.3e} barrier cpu {Ec - E0:.9f} gpu {Eg - E0:.9f} diff {Eg - Ec:.2e}")
    sR.close()

probe(4, 8)
probe(6, 32)
if len(sys.argv) > 1 and sys.argv[1] == "big":
    probe(20, 64, a=2.8553)
