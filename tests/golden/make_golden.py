"""Writes tests/golden/*.npz from the CPU oracle (canonical build, -ffp-contract=off).

The reference itself cannot be built in this image, so its own known-answer constants are
transcribed, with file:line, in tests/test_oracle_goldens.py and reference_kat.json; the vectors written HERE
are oracle outputs on the BASELINE configurations, committed so that (a) the oracle is checked for drift on CPU
and (b) the CUDA path is checked against fixed files on the GPU box, independent of the oracle build there.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.common import WEN, bcc_vacancy, pair_multiset  # noqa: E402


def neighbour_digest(counts, idx, sh):
    """One 64-bit digest per atom of its sorted (real index, shift code) multiset."""
    out = np.zeros(len(counts), dtype=np.uint64)
    for i, pairs in enumerate(pair_multiset(counts, idx, sh)):
        h = hashlib.blake2b(np.asarray(pairs, dtype=np.int64).tobytes(), digest_size=8).digest()
        out[i] = np.frombuffer(h, dtype=np.uint64)[0]
    return out


def c1():
    """C1: bcc Fe 10^3 (a = 2.8553) - atom 1, jitter N(0, 0.05) Xoshiro{1,2,3,4}; rebuild(5.3) + energy + gradient."""
    pos, ex, ids = bcc_vacancy(10, jitter=0.05, seed=(1, 2, 3, 4))
    tab = O.Tables(WEN)
    s = O.System(tab, ex)
    s.rebuild(tab.rcut, pos, 4)
    counts, idx, sh = s.export()
    fz = np.zeros(len(ids), np.uint8)
    np.savez_compressed(os.path.join(HERE, "c1_oracle.npz"), pos=pos, extents=ex, ids=ids, cells=s.cells(), counts=counts,
                        digest=neighbour_digest(counts, idx, sh), energy=s.energy(ids, 4), gradient=s.gradient(ids, fz, 4))


def c3():
    """C3-style 6^3 Fe-H cell (431 atoms): relaxed minimum, Hessian eigenvalue spectrum digest, four dimer searches."""
    base, ex, ids = bcc_vacancy(6, a=2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    tab = O.Tables(WEN)
    s = O.System(tab, ex)
    r0, g0, failed, cnt = s.minimise(base, ids, fz, threads=8)
    assert not failed
    s.rebuild(tab.rcut, r0, 4)
    e0 = s.energy(ids, 4)
    H = s.hessian(ids, fz, 4, mass_weight=True)
    ev = np.linalg.eigvalsh(np.tril(H) + np.tril(H, -1).T)
    starts = [s.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(4)]
    P, A = np.stack([x[0] for x in starts]), np.stack([x[1] for x in starts])
    o_r, o_ax, code, c = s.find_sp_batch(P, A, ids, fz, r0, threads=4)
    bar = []
    for k in range(4):
        s.rebuild(tab.rcut, o_r[k], 4)
        bar.append(s.energy(ids, 4) - e0)
    np.savez_compressed(os.path.join(HERE, "c3_oracle.npz"), base=base, extents=ex, ids=ids, minimum=r0, energy_min=e0,
                        eigenvalues=ev, start_pos=P, start_axis=A, exit=code, saddle=o_r, barrier=np.array(bar), evals=c[:, 0])


if __name__ == "__main__":
    O.build()
    c1()
    c3()
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])
