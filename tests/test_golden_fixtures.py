"""Committed fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the canonical oracle build).
CPU: the oracle still reproduces them (drift guard). GPU: the CUDA path reproduces them through the C ABI, with the
north-star tolerances: cell ids and neighbour sets bit-exact, energy 1e-10 relative, forces 1e-9 eV/A,
barriers 1e-6 eV, mass-weighted Hessian eigenvalues (-> prefactors) 1e-4 relative."""
import os

import numpy as np
import pytest

from tests.common import ROOT, WEN
from tests.golden.make_golden import neighbour_digest

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def c1():
    return np.load(os.path.join(GOLD, "c1_oracle.npz"))


@pytest.fixture(scope="module")
def c3():
    return np.load(os.path.join(GOLD, "c3_oracle.npz"))


def test_oracle_reproduces_c1(oracle_lib, c1):
    tab = oracle_lib.Tables(WEN)
    s = oracle_lib.System(tab, c1["extents"])
    s.rebuild(tab.rcut, c1["pos"], 2)
    counts, idx, sh = s.export()
    assert np.array_equal(s.cells(), c1["cells"]) and np.array_equal(counts, c1["counts"])
    assert np.array_equal(neighbour_digest(counts, idx, sh), c1["digest"])
    assert abs(s.energy(c1["ids"], 2) - float(c1["energy"])) <= 1e-13 * abs(float(c1["energy"]))  # OpenMP reduction order
    assert np.array_equal(s.gradient(c1["ids"], np.zeros(len(c1["ids"]), np.uint8), 1), c1["gradient"])


def test_oracle_reproduces_c3_minimum_and_one_search(oracle_lib, c3):
    tab = oracle_lib.Tables(WEN)
    s = oracle_lib.System(tab, c3["extents"])
    fz = np.zeros(len(c3["ids"]), np.uint8)
    r0, _, failed, _ = s.minimise(c3["base"], c3["ids"], fz, threads=2)
    assert not failed and np.array_equal(r0, c3["minimum"])
    k = int(np.argmin(c3["evals"]))  # the cheapest of the four searches keeps the CPU suite short
    o_r, o_ax, code, cnt = s.find_sp(c3["start_pos"][k], c3["start_axis"][k], c3["ids"], fz, c3["minimum"])
    assert code == c3["exit"][k] and cnt[0] == c3["evals"][k] and np.array_equal(o_r, c3["saddle"][k])


@pytest.fixture(scope="module")
def gpu():
    from openfly_b200 import fly

    ctx = fly.Context(0)
    data = fly.DataEAM(ctx, WEN)
    yield fly, ctx, data
    data.close()
    ctx.close()


@pytest.mark.gpu
def test_gpu_reproduces_c1(gpu, c1):
    fly, ctx, data = gpu
    s = fly.System(data, fly.Box(tuple(c1["extents"])), c1["ids"])
    nl = fly.List(s, data.rcut())
    nl.rebuild(c1["pos"])
    counts, idx, sh = nl.export()
    assert np.array_equal(nl.cells(), c1["cells"])            # bit-exact cell assignment
    assert np.array_equal(counts, c1["counts"])
    assert np.array_equal(neighbour_digest(counts, idx, sh), c1["digest"])  # bit-exact neighbour sets
    pot = fly.Generic(data)
    E = pot.energy(nl)[0]
    assert abs(E - float(c1["energy"])) <= 1e-10 * abs(float(c1["energy"]))
    assert np.abs(pot.gradient(nl)[0] - c1["gradient"]).max() < 1e-9
    s.close()


@pytest.mark.gpu
def test_gpu_reproduces_c3(gpu, c3):
    fly, ctx, data = gpu
    ids = c3["ids"]
    pot = fly.Generic(data)
    s = fly.System(data, fly.Box(tuple(c3["extents"])), ids)
    r1, g1, failed, cnt = fly.LBFGS().minimise(s, c3["base"], pot)
    assert not failed[0] and np.abs(r1[0] - c3["minimum"]).max() < 1e-6
    nl = fly.List(s, data.rcut())
    nl.rebuild(c3["minimum"])
    e0 = pot.energy(nl)[0]
    assert abs(e0 - float(c3["energy_min"])) <= 1e-10 * abs(e0)
    H = pot.hessian(nl, mass_weight=True)
    ev = np.linalg.eigvalsh(np.tril(H) + np.tril(H, -1).T)
    big = np.abs(c3["eigenvalues"]) > 1e-6  # the modes that enter the prefactor (find.cpp:852-864)
    assert np.abs(ev[big] / c3["eigenvalues"][big] - 1).max() < 1e-4
    assert np.abs(ev[~big]).max() < 1e-6
    s.close()

    R = len(c3["exit"])
    sR = fly.System(data, fly.Box(tuple(c3["extents"])), ids, replicas=R)
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(sR, c3["start_pos"], c3["start_axis"], c3["minimum"], pot)
    assert code.tolist() == c3["exit"].tolist()
    nlR = fly.List(sR, data.rcut())
    nlR.rebuild(out_r)
    bar = pot.energy(nlR) - e0
    for k in range(R):
        if code[k] == 0:
            assert abs(bar[k] - c3["barrier"][k]) < 1e-6
            assert np.abs(out_r[k] - c3["saddle"][k]).max() < 1e-5
    sR.close()
