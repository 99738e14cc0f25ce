"""Shared builders for the tests: the synthetic systems of SURVEY.md 8(d), made with the oracle's
input builders (oracle = test infrastructure, allowed here)."""
import os

import numpy as np

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEN = os.path.join(ROOT, "data", "wen.eam.fs")


def golden_cell(erase=True):
    """tests/potential/EAM/eam.cpp:36-96: 7x7x7 bcc Fe at the 300 K lattice parameter, atom #1 erased."""
    T = 300.0
    V = 11.64012 + T * (9.37798e-5 + T * (3.643134e-7 + T * (1.851593e-10 + T * 5.669148e-14)))
    a = (2 * V) ** (1.0 / 3)
    lat = []
    for k in range(7):
        for j in range(7):
            for i in range(7):
                for off in (0.0, 0.5):
                    lat.append(((i + off) * a, (j + off) * a, (k + off) * a))
    lat = np.array(lat)
    if erase:
        lat = np.delete(lat, 1, axis=0)
    return lat, np.array([7 * a] * 3), np.zeros(len(lat), dtype=np.uint32)


def bcc_vacancy(n, a=2.8553, remove=(1,), jitter=0.0, seed=(1, 2, 3, 4), hydrogen=False):
    """C1/C2/C3-style cells: motif_to_lattice bcc n^3, remove atoms, optional H, optional jitter."""
    pos = O.bcc_lattice(a, n, n, n)
    pos = np.delete(pos, list(remove), axis=0)
    ids = np.zeros(len(pos), dtype=np.uint32)
    if hydrogen:  # apps/hydrogen.cpp:137-139
        pos = np.vstack([pos, [[2.857 / 2 + 3.14, 2.857 / 2 + 3.14, 2.857 / 4 + 3.14]]])
        ids = np.append(ids, np.uint32(1))
    if jitter > 0:
        pos = O.jitter(pos, seed, jitter)
    return np.ascontiguousarray(pos), np.array([n * a] * 3), ids


def pair_multiset(counts, idx, sh):
    """Per-atom sorted list of (real index, shift code) -- order-independent neighbour sets."""
    out = []
    off = 0
    for c in counts:
        pairs = sorted(zip(idx[off:off + c].tolist(), sh[off:off + c].tolist()))
        out.append(pairs)
        off += c
    return out
