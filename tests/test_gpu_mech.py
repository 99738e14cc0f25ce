"""GPU parity of the batched post-search numerics (ofly_mech_*: do_adj_min, saddle_2_mech's energies / Hessian / eigenvalue
rules, calc_minima_hess, prefactor -- src/saddle/find.cpp) against the CPU composition in tests/mech_oracle.py, on the real
apps/hydrogen-sized cell (431 atoms). Tolerances are BASELINE.json's: barriers 1e-6 eV, prefactors 1e-4 relative."""
import numpy as np
import pytest

from tests import mech_oracle as MO
from tests.common import WEN, bcc_vacancy, golden_cell

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from openfly_b200 import fly

    ctx = fly.Context(0)
    data = fly.DataEAM(ctx, WEN)
    yield fly, ctx, data
    data.close()
    ctx.close()


@pytest.fixture(scope="module")
def c3(gpu, oracle_lib):
    """The C3 cell, its minimum and the saddles of 64 searches (GPU), shared by the tests below."""
    fly, ctx, data = gpu
    otab = oracle_lib.Tables(WEN)
    base, ex, ids = bcc_vacancy(6, a=2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    osys = oracle_lib.System(otab, ex)
    r0, _, failed, _ = osys.minimise(base, ids, fz, threads=8)
    assert not failed
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(64)]
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    sR = fly.System(data, fly.Box(tuple(ex)), ids, replicas=64)
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(sR, P, A, r0, fly.Generic(data))
    sR.close()
    ok = np.nonzero(code == 0)[0]
    assert len(ok) >= 20
    return otab, osys, ex, ids, fz, r0, out_r[ok], out_ax[ok]


def test_batched_hessians_equal_single_and_are_reproducible(gpu, oracle_lib):
    """One launch for every replica (no atomics, fixed sum order): equal to the per-replica call bit for bit, twice the same bits,
    and 1e-9 against the oracle."""
    fly, ctx, data = gpu
    pos, ex, ids = bcc_vacancy(4, a=2.855, remove=(1,), jitter=0.04, hydrogen=True)
    R = 3
    P = np.stack([pos, pos + 0.01 * np.sin(np.arange(pos.size).reshape(pos.shape)), pos[::-1][::-1] * 1.0])
    fz = np.zeros(len(ids), np.uint8)
    fz[7] = 1
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=R)
    nl = fly.List(s, data.rcut())
    nl.rebuild(P)
    pot = fly.Generic(data)
    Hb = pot.hessian_batch(nl)
    assert np.array_equal(Hb, pot.hessian_batch(nl))
    for r in range(R):
        assert np.array_equal(np.tril(Hb[r]), np.tril(pot.hessian(nl, replica=r)))
    otab = oracle_lib.Tables(WEN)
    osys = oracle_lib.System(otab, ex)
    osys.rebuild(otab.rcut, P[1], 4)
    Ho = osys.hessian(ids, fz, 4)
    assert np.abs(np.tril(Hb[1]) - np.tril(Ho)).max() < 1e-9
    assert np.all(Hb[1][21:24, :] == 0) and np.all(Hb[1][:, 21:24] == 0)  # frozen atom: zero rows / columns
    s.close()


def test_minimum_rules_and_sep_list(gpu, c3):
    fly, ctx, data = gpu
    otab, osys, ex, ids, fz, r0, sps, axs = c3
    mech = fly.Mechanisms(data, fly.Box(tuple(ex)), ids, replicas=8)
    lp, nz, nf, eig = mech.minimum(r0)
    lp_o, nz_o, eig_o = MO.minimum(osys, otab, ids, fz, r0)
    assert (nz, nf) == (nz_o, 0) and nz == 3
    assert abs(lp - lp_o) < 1e-7 * abs(lp_o) and np.abs(eig - eig_o).max() < 1e-9
    assert np.array_equal(mech.sep_list(r0), MO.sep_list(r0, ex))
    # a saddle is not a minimum: the reference's verify fires (find.cpp:1120)
    with pytest.raises(fly.OflyError, match="has eigenvalue"):
        mech.minimum(sps[0])
    mech.close()


def test_every_saddle_prefactor_and_barrier_against_oracle(gpu, c3, capsys):
    """Every saddle of the 64-search C3 batch: barrier 1e-6 eV, prefactor 1e-4 relative, same validity verdict."""
    fly, ctx, data = gpu
    otab, osys, ex, ids, fz, r0, sps, axs = c3
    mech = fly.Mechanisms(data, fly.Box(tuple(ex)), ids, replicas=16)  # fewer slots than saddles: chunked
    res = mech.prefactors(r0, sps)
    lp_min, nz_min, _ = MO.minimum(osys, otab, ids, fz, r0)
    osys.rebuild(otab.rcut, r0, 8)
    e_min = osys.energy(ids, 8)
    n_valid = 0
    for k in range(len(sps)):
        osys.rebuild(otab.rcut, sps[k], 8)
        e_sp = osys.energy(ids, 8)
        freq = MO.eigenvalues(osys.hessian(ids, fz, 8, mass_weight=True))
        assert abs(res["barrier"][k] - (e_sp - e_min)) < 1e-6
        assert abs(res["min_mode"][k] - freq[0]) < 1e-8
        rest = freq[1:]
        valid_o = freq[0] < -1e-6 and not (rest < -1e-6).any() and int((np.abs(rest) <= 1e-6).sum()) == nz_min
        assert bool(res["valid"][k]) == bool(valid_o), (k, res["status"][k], freq[:6])
        if valid_o:
            nu_o = MO.prefactor(lp_min, float(np.log(rest[rest > 1e-6]).sum()))
            assert abs(res["prefactor"][k] / nu_o - 1) < 1e-4
            assert 1e11 < res["prefactor"][k] < 1e16
            n_valid += 1
    with capsys.disabled():
        print(f"\n[prefactor parity] {n_valid}/{len(sps)} saddles valid first-order, all prefactors within 1e-4 of the oracle; "
              f"statuses {sorted(set(res['status'].tolist()))}")
    assert n_valid >= 15
    mech.close()


def test_saddle_to_mechanism_pathways_against_oracle(gpu, c3, capsys):
    """do_adj_min + saddle_2_mech for the saddles of the batch at once vs the CPU composition one by one: same verdicts, pathway
    end points to 1e-5 A, barrier / delta to 1e-6 eV, prefactor to 1e-4."""
    fly, ctx, data = gpu
    otab, osys, ex, ids, fz, r0, sps, axs = c3
    S = min(len(sps), 12)
    centre = len(ids) - 1
    mech = fly.Mechanisms(data, fly.Box(tuple(ex)), ids, replicas=10)
    lp, nz, nf, _ = mech.minimum(r0)
    freeze = int(mech.sep_list(r0)[centre])
    res = mech.from_saddles(r0, sps[:S], axs[:S], centre=centre, freeze_atom=freeze)
    lp_o, nz_o, _ = MO.minimum(osys, otab, ids, fz, r0)
    n_ok, verdicts = 0, []
    for k in range(S):
        o = MO.saddle_to_mech(osys, otab, ids, fz, r0, sps[k], axs[k], centre, freeze, nz_o)
        verdicts.append((int(res["status"][k]), o["status"]))
        assert res["status"][k] == o["status"], (k, verdicts)
        if o["status"] in (1,):
            continue
        for name in ("rev", "sp", "fwd"):
            assert np.abs(res[name][k] - o[name]).max() < 1e-5
        if o["status"] == 2:
            continue
        for name in ("i2r", "r2f", "r2w", "w2f"):
            assert abs(res[name][k] - o[name]) < 1e-5
        if o["status"] == 0:
            assert abs(res["barrier"][k] - o["barrier"]) < 1e-6 and abs(res["delta"][k] - o["delta"]) < 1e-6
            assert abs(res["prefactor"][k] / MO.prefactor(lp_o, o["log_prod_eigen"]) - 1) < 1e-4
            n_ok += 1
    with capsys.disabled():
        print(f"\n[mechanism parity] {n_ok}/{S} saddles became mechanisms on both sides; (gpu, oracle) verdicts {verdicts}")
    assert n_ok >= 3
    mech.close()


def test_own_eigenvalue_kernel_against_cusolver_and_numpy(gpu, c3, oracle_lib):
    """eigen.cuh (Householder tridiagonalisation with lazy rank-2 updates + Sturm bisection, one CTA per matrix) against
    cuSOLVER's Dsyevd and numpy on the same mass-weighted Hessians: 1e-9 absolute on every eigenvalue, including the three
    zero modes and the negative mode of a saddle; frozen atoms (zero rows / columns) included."""
    import os

    fly, ctx, data = gpu
    otab, osys, ex, ids, fz, r0, sps, axs = c3
    pot = fly.Generic(data)
    s1 = fly.System(data, fly.Box(tuple(ex)), ids)
    nl = fly.List(s1, data.rcut())
    got = {}
    for mode in ("own", "cusolver"):
        os.environ["OFLY_EIGEN"] = mode
        try:
            nl.rebuild(sps[0])
            got[mode] = pot.hessian_eigenvalues(nl, mass_weight=True)
        finally:
            del os.environ["OFLY_EIGEN"]
    s1.close()
    osys.rebuild(otab.rcut, sps[0], 8)
    ref = MO.eigenvalues(osys.hessian(ids, fz, 8, mass_weight=True))
    assert np.abs(got["own"] - ref).max() < 1e-9 and np.abs(got["cusolver"] - ref).max() < 1e-9
    assert got["own"][0] < -1e-3 and np.abs(got["own"][1:4]).max() < 1e-8 and np.all(np.diff(got["own"]) >= 0)
    # a batch (every matrix in flight at once) gives the same numbers as one at a time, and frozen atoms are handled
    fz2 = fz.copy()
    fz2[[3, 40]] = 1
    mech = fly.Mechanisms(data, fly.Box(tuple(ex)), ids, fz2, replicas=8)
    os.environ["OFLY_EIGEN"] = "own"
    try:
        many = mech.prefactors(r0, sps[:8], opt=fly.mech_options(num_zero_modes=6))
    finally:
        del os.environ["OFLY_EIGEN"]
    os.environ["OFLY_EIGEN"] = "cusolver"
    try:
        lib = mech.prefactors(r0, sps[:8], opt=fly.mech_options(num_zero_modes=6))
    finally:
        del os.environ["OFLY_EIGEN"]
    mech.close()
    assert np.array_equal(many["status"], lib["status"])
    assert np.abs(many["min_mode"] - lib["min_mode"]).max() < 1e-9
    ok = many["status"] == 0
    assert np.abs(many["log_prod_eigen"][ok] - lib["log_prod_eigen"][ok]).max() < 1e-7


def test_for_neighbours_of_one_atom_on_device(gpu, c3):
    """ofly_neigh_for_atom against the oracle's List::for_neighbours (list.hpp:138-154): same neighbours, r and dr."""
    fly, ctx, data = gpu
    otab, osys, ex, ids, fz, r0, sps, axs = c3
    s = fly.System(data, fly.Box(tuple(ex)), ids)
    nl = fly.List(s, 4.0 + 0.3)
    nl.rebuild(sps[0])
    osys.rebuild(4.3, sps[0])
    for atom in (0, 17, len(ids) - 1):
        for r_cut in (4.0, 2.6):
            idx, r, dr = nl.neighbours_of(atom, r_cut)
            o_idx, o_r, o_dr = osys.for_neighbours(atom, r_cut)
            a, b = np.lexsort((r, idx)), np.lexsort((o_r, o_idx))
            assert np.array_equal(idx[a], np.asarray(o_idx)[b])
            assert np.abs(r[a] - np.asarray(o_r)[b]).max() < 1e-12 and np.abs(dr[a] - np.asarray(o_dr)[b]).max() < 1e-12
    seen = []
    nl.for_neighbours(5, 3.0, lambda n, r, dr: seen.append((n, r)))
    assert len(seen) == len(osys.for_neighbours(5, 3.0)[0])
    with pytest.raises(fly.OflyError):
        nl.neighbours_of(0, 5.0)  # beyond the list radius (list.hpp:141-143)
    s.close()


def test_master_perturb_on_device(gpu, c3):
    """Master::perturb (find.cpp:1147-1214) with the caller's normal draws against a numpy restatement: atoms within r_pert of
    the centre get stddev * z * (1 - r / r_pert) and an axis component, the centre a full kick, the axis is normalised."""
    fly, ctx, data = gpu
    otab, osys, ex, ids, fz, r0, sps, axs = c3
    N, centre, r_pert = len(ids), len(ids) - 1, 4.0
    rng = np.random.default_rng(3)
    S = 5
    sd = np.abs(rng.normal(0.6, 0.2, S))
    z = rng.normal(size=(S, N, 6))
    mech = fly.Mechanisms(data, fly.Box(tuple(ex)), ids, replicas=2)
    pos, ax = mech.perturb(r0, centre, sd, z, r_pert)
    mech.close()
    osys.rebuild(r_pert, r0)
    o_idx, o_r, _ = osys.for_neighbours(centre, r_pert)
    for s in range(S):
        want_p, want_a = r0.copy(), np.zeros_like(r0)
        for n, r in zip(o_idx, o_r):
            want_p[n] += sd[s] * z[s, n, :3] * (1 - r / r_pert)
            want_a[n] += z[s, n, 3:]
        want_p[centre] += sd[s] * z[s, centre, :3]
        want_a[centre] += z[s, centre, 3:]
        want_a /= np.sqrt((want_a ** 2).sum())
        assert np.abs(pos[s] - want_p).max() < 1e-12 and np.abs(ax[s] - want_a).max() < 1e-12
        assert abs((ax[s] ** 2).sum() - 1) < 1e-12
        moved = np.nonzero(np.abs(pos[s] - r0).max(axis=1) > 0)[0]
        assert set(moved.tolist()) == set(int(n) for n in o_idx) | {centre}


def test_skmc_relax_step(gpu, c3, oracle_lib):
    """skmc.hpp:297-309 in one call against the CPU composition: E0, relaxed state, aligned displacement, Ef."""
    fly, ctx, data = gpu
    otab, osys, ex, ids, fz, r0, sps, axs = c3
    raw = oracle_lib.jitter(r0, (7, 7, 7, 7), 0.03)
    mech = fly.Mechanisms(data, fly.Box(tuple(ex)), ids, replicas=1)
    got = mech.skmc_relax_step(r0, raw)
    mech.close()
    rel, _, failed, _ = osys.minimise(raw, ids, fz, threads=8)
    rel = rel + (raw.mean(axis=0) - rel.mean(axis=0))
    osys.rebuild(otab.rcut, r0)
    e0 = osys.energy(ids)
    osys.rebuild(otab.rcut, rel)
    ef = osys.energy(ids)
    assert got["failed"] == bool(failed)
    assert abs(got["E0"] - e0) < 1e-9 and abs(got["Ef"] - ef) < 1e-8
    assert np.abs(got["rel_recon"] - rel).max() < 1e-6
    assert abs(got["dR"] - np.sqrt(((rel - raw) ** 2).sum())) < 1e-5


def test_neighbourhoods_of_every_atom_in_one_pass(gpu, c3):
    """ofly_neigh_within_all (the loop of Catalogue::rebuild_impl, src/env/catalogue.cpp:80-96, at r_env): CSR of every atom's
    neighbours within r_cut against the oracle's for_neighbours atom by atom."""
    fly, ctx, data = gpu
    otab, osys, ex, ids, fz, r0, sps, axs = c3
    s = fly.System(data, fly.Box(tuple(ex)), ids)
    nl = fly.List(s, 5.2)
    nl.rebuild(sps[1])
    osys.rebuild(5.2, sps[1])
    off, idx, r, dr = nl.neighbours_within(5.2)
    assert off[0] == 0 and off[-1] == len(idx)
    for atom in list(range(0, len(ids), 37)) + [len(ids) - 1]:
        o_idx, o_r, o_dr = osys.for_neighbours(atom, 5.2)
        a = slice(off[atom], off[atom + 1])
        p, q = np.lexsort((r[a], idx[a])), np.lexsort((o_r, o_idx))
        assert np.array_equal(idx[a][p], np.asarray(o_idx)[q])
        assert np.abs(r[a][p] - np.asarray(o_r)[q]).max() < 1e-12 and np.abs(dr[a][p] - np.asarray(o_dr)[q]).max() < 1e-12
        one = nl.neighbours_of(atom, 5.2)
        assert np.array_equal(one[0], idx[a]) and np.array_equal(one[1], r[a])  # same order as the single-atom call
    off2, idx2, _, _ = nl.neighbours_within(3.0)
    assert off2[-1] < off[-1]
    s.close()


def test_python_list_refuses_stale_device_state(gpu, c3):
    """fly.py's System holds ONE device list (fly.hpp likewise; libfly.hpp owns one per List): a List whose device state was
    replaced by another List or by a search must not be evaluated silently on the wrong positions."""
    fly, ctx, data = gpu
    otab, osys, ex, ids, fz, r0, sps, axs = c3
    s = fly.System(data, fly.Box(tuple(ex)), ids)
    pot = fly.Generic(data)
    a, b = fly.List(s, data.rcut()), fly.List(s, data.rcut() + 0.4)
    a.rebuild(r0)
    e = pot.energy(a)[0]
    b.rebuild(sps[0])
    with pytest.raises(fly.OflyError, match="replaced"):
        pot.energy(a)
    a.rebuild(r0)
    assert pot.energy(a)[0] == e
    fly.LBFGS().minimise(s, sps[0], pot)
    with pytest.raises(fly.OflyError, match="replaced"):
        pot.gradient(a)
    s.close()
