"""GPU parity: minimise::LBFGS::minimise and saddle::Dimer::find_sp (lock-step engine, through the C ABI)
against the CPU oracle on identical start states.

The reference has no tests for saddle/ or minimise/ (SURVEY.md 4), so these trajectories are pinned
oracle-vs-GPU only: same exit codes, saddle positions, barriers to 1e-6 eV (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from tests.common import WEN, bcc_vacancy

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
def otab(oracle_lib):
    return oracle_lib.Tables(WEN)


@pytest.fixture(scope="module")
def small_cell(gpu, otab, oracle_lib):
    """4^3 bcc Fe - 1 vacancy + 1 H (128 atoms), relaxed by the oracle; 8 Master::perturb start states."""
    base, ex, ids = bcc_vacancy(4, a=2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    osys = oracle_lib.System(otab, ex)
    r0, g0, failed, cnt = osys.minimise(base, ids, fz, threads=4)
    assert not failed
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(8)]
    P = np.stack([s[0] for s in starts])
    A = np.stack([s[1] for s in starts])
    return base, ex, ids, fz, osys, r0, P, A


def _energy(osys, otab, ids, r):
    osys.rebuild(otab.rcut, r)
    return osys.energy(ids, 4)


def test_minimise_matches_oracle(gpu, otab, small_cell):
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    s = fly.System(data, fly.Box(tuple(ex)), ids)
    r1, g1, failed, cnt = fly.LBFGS().minimise(s, base, fly.Generic(data))
    assert not failed[0]
    assert np.abs(r1[0] - r0).max() < 1e-6
    assert np.sqrt((g1[0] ** 2).sum()) < 1e-6  # f2norm
    E0, E1 = _energy(osys, otab, ids, r0), _energy(osys, otab, ids, r1[0])
    assert abs(E1 - E0) < 1e-9
    assert cnt[0, 0] >= 2 and cnt[0, 1] >= 1
    s.close()


def test_minimise_batch_iter_max_and_frozen(gpu, otab, oracle_lib, small_cell):
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    fz2 = fz.copy()
    fz2[3] = 1
    R = 3
    pos = np.stack([oracle_lib.jitter(base, (9, 2, 3, 4 + r), 0.05) for r in range(R)])
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz2, replicas=R)
    out_r, out_g, failed, cnt = fly.LBFGS().minimise(s, pos, fly.Generic(data))
    o_r, o_g, o_fail, o_cnt = osys.minimise_batch(pos, ids, fz2, threads=3)
    assert failed.tolist() == [bool(x) for x in o_fail]
    assert np.abs(out_r - o_r).max() < 1e-6
    assert np.all(out_r[:, 3] == pos[:, 3])  # frozen atom never moves
    assert np.all(out_g[:, 3] == 0)
    # iteration limit: returns "failed" with the state reached
    lim = fly.LBFGS(fly.lbfgs_options(iter_max=3))
    out_r, out_g, failed, cnt = lim.minimise(s, pos, fly.Generic(data))
    assert failed.all() and (cnt[:, 3] == 3).all() and (cnt[:, 0] == 4).all()
    iopt = oracle_lib.LBFGS_IOPT.copy()
    iopt[1] = 3
    o_r, o_g, o_fail, o_cnt = osys.minimise_batch(pos, ids, fz2, threads=3, iopt=iopt)
    assert np.abs(out_r - o_r).max() < 1e-9
    s.close()


def _check_dimer_against_oracle(fly, data, otab, ids, fz, ex, osys, r0, P, A, threads=8, **kw):
    R = len(P)
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=R)
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(s, P, A, r0, fly.Generic(data), **kw)
    o_r, o_ax, o_code, o_cnt = osys.find_sp_batch(P, A, ids, fz, r0, threads=threads, **kw)
    s.close()
    assert code.tolist() == o_code.tolist()
    E0 = _energy(osys, otab, ids, r0)
    n_ok = 0
    for k in range(R):
        if code[k] == 0:
            n_ok += 1
            assert np.abs(out_r[k] - o_r[k]).max() < 1e-5
            assert abs(abs((out_ax[k] * o_ax[k]).sum()) - 1) < 1e-3  # same unstable mode (sign is free; rotor theta_tol = 0.5 deg)
            barrier_g = _energy(osys, otab, ids, out_r[k]) - E0
            barrier_o = _energy(osys, otab, ids, o_r[k]) - E0
            assert abs(barrier_g - barrier_o) < 1e-6  # north_star tolerance on saddle barriers
            assert barrier_g > 0
    return code, cnt, o_cnt, n_ok


def test_dimer_matches_oracle_small(gpu, otab, small_cell):
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    code, cnt, o_cnt, n_ok = _check_dimer_against_oracle(fly, data, otab, ids, fz, ex, osys, r0, P, A)
    assert n_ok >= 1
    # work counters: same order of magnitude as the CPU path (trajectories agree until rounding separates them)
    assert np.all(cnt[:, 0] > 0)
    assert abs(int(cnt[:, 0].sum()) - int(o_cnt[:, 0].sum())) <= 0.2 * o_cnt[:, 0].sum()


def test_dimer_c3_cell(gpu, otab, oracle_lib):
    """apps/hydrogen-sized cell (6^3, 431 atoms, L ~ 3.1 list radii: many periodic images)."""
    fly, ctx, data = gpu
    base, ex, ids = bcc_vacancy(6, a=2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    osys = oracle_lib.System(otab, ex)
    r0, _, failed, _ = osys.minimise(base, ids, fz, threads=8)
    assert not failed
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(12)]
    P = np.stack([s[0] for s in starts])
    A = np.stack([s[1] for s in starts])
    code, cnt, o_cnt, n_ok = _check_dimer_against_oracle(fly, data, otab, ids, fz, ex, osys, r0, P, A)
    assert n_ok >= 2


def test_dimer_history_collision_and_limits(gpu, otab, small_cell):
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    pot = fly.Generic(data)
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=len(P))
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(s, P, A, r0, pot)
    found = [k for k in range(len(P)) if code[k] == 0]
    assert found
    # cache = the first found saddle; a wide collision cone sends every search that heads there home early
    hist = out_r[found[0]][None]
    theta = np.deg2rad(60.0)
    r2, a2, code2, cnt2 = fly.Dimer().find_sp(s, P, A, r0, pot, hist_sp=hist, theta_tol=theta)
    o_r, o_ax, o_code, o_cnt = osys.find_sp_batch(P, A, ids, fz, r0, hist_sp=hist, theta_tol=theta, threads=8)
    assert code2.tolist() == o_code.tolist()
    assert code2[found[0]] == fly.Dimer.collision
    # step limit
    lim = fly.Dimer(fly.dimer_options(max_steps=2))
    r3, a3, code3, cnt3 = lim.find_sp(s, P, A, r0, pot)
    assert (code3 == fly.Dimer.iter_max).all() and (cnt3[:, 3] == 2).all()
    s.close()


def test_dimer_frozen_no_projection(gpu, otab, small_cell):
    """count_frozen > 0 switches trans_proj off (rotor.cpp:92-94,159-161); find.cpp:116 calls it that way."""
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    fz2 = fz.copy()
    fz2[0] = 1
    A2 = A.copy()
    A2[:, 0] = 0
    A2 /= np.sqrt((A2 ** 2).sum(axis=(1, 2)))[:, None, None]
    P2 = P.copy()
    P2[:, 0] = r0[0]
    _check_dimer_against_oracle(fly, data, otab, ids, fz2, ex, osys, r0, P2[:4], A2[:4], threads=4, count_frozen=1)


def test_host_loop_equals_device_loop(gpu, small_cell):
    """OFLY_SEARCH_LOOP=host drives the same tick kernels from the host: bit-identical results."""
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    pot = fly.Generic(data)
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=len(P))
    ref = fly.Dimer().find_sp(s, P, A, r0, pot)
    os.environ["OFLY_SEARCH_LOOP"] = "host"
    try:
        alt = fly.Dimer().find_sp(s, P, A, r0, pot)
    finally:
        del os.environ["OFLY_SEARCH_LOOP"]
    for a, b in zip(ref, alt):
        assert np.array_equal(a, b)
    s.close()


def test_single_pass_lbfgs_equals_two_loop(gpu, otab, small_cell):
    """The optional two-pass Gram L-BFGS step (OFLY_LBFGS=gram, search.cuh newton_step_gram) against the default, the literal
    pass-per-pair two-loop recursion of lbfgs_core.cpp:29-87: the two round alpha/beta differently, so trajectories agree to
    rounding: same exits, same saddles."""
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    pot = fly.Generic(data)
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=len(P))
    old = fly.Dimer().find_sp(s, P, A, r0, pot)
    os.environ["OFLY_LBFGS"] = "gram"
    try:
        new = fly.Dimer().find_sp(s, P, A, r0, pot)
    finally:
        del os.environ["OFLY_LBFGS"]
    assert new[2].tolist() == old[2].tolist()
    for k in range(len(P)):
        if new[2][k] == 0:
            assert np.abs(new[0][k] - old[0][k]).max() < 1e-5
    assert abs(int(new[3][:, 0].sum()) - int(old[3][:, 0].sum())) <= 0.2 * old[3][:, 0].sum()
    s.close()


def test_rebuild_if_node_equals_flat_graph(gpu, small_cell):
    """OFLY_REBUILD_IF=0 keeps the five rebuild kernels unconditional in every tick: bit-identical results."""
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    pot = fly.Generic(data)
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=len(P))
    ref = fly.Dimer().find_sp(s, P, A, r0, pot)
    os.environ["OFLY_REBUILD_IF"] = "0"
    try:
        alt = fly.Dimer().find_sp(s, P, A, r0, pot)
    finally:
        del os.environ["OFLY_REBUILD_IF"]
    for a, b in zip(ref, alt):
        assert np.array_equal(a, b)
    s.close()


def test_out_of_bounds_reported_per_replica(gpu, small_cell):
    """A replica whose atoms overlap overruns the F(rho) table: status -5 (the reference throws
    Spline::OutOfBounds, which Master catches and retries, find.cpp:718-722); the others are unaffected."""
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    bad = P[:3].copy()
    bad[1, 1:40] = bad[1, 0] + np.linspace(0.01, 0.4, 39)[:, None]
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=3)
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(s, bad, A[:3], r0, fly.Generic(data))
    assert code[1] == -5 and code[0] != -5 and code[2] != -5
    with pytest.raises(fly.OutOfBounds):
        fly.Dimer().find_sp(s, bad, A[:3], r0, fly.Generic(data), raise_oob=True)
    s.close()


def test_barriers_and_vineyard_prefactors_c3(gpu, otab, oracle_lib):
    """C3's deliverable: for every saddle found, barrier (1e-6 eV) and harmonic prefactor (1e-4 relative, BASELINE north
    star) from the GPU's mass-weighted Hessians against the same quantities from the oracle's Hessians.
    find.cpp:284-290,846-871,1105-1135."""
    fly, ctx, data = gpu
    base, ex, ids = bcc_vacancy(5, a=2.855, remove=(1,), hydrogen=True)  # 250 atoms: the dense eigensolves stay cheap
    fz = np.zeros(len(ids), np.uint8)
    osys = oracle_lib.System(otab, ex)
    r0, _, failed, _ = osys.minimise(base, ids, fz, threads=8)
    assert not failed
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(16)]
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    pot = fly.Generic(data)
    sR = fly.System(data, fly.Box(tuple(ex)), ids, replicas=len(P))
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(sR, P, A, r0, pot)
    sR.close()
    o_r, o_ax, o_code, o_cnt = osys.find_sp_batch(P, A, ids, fz, r0, threads=8)
    # searches that reached the same saddle on both sides (a long trajectory may branch out, see the 64-search test)
    both = [k for k in range(len(P)) if code[k] == 0 and o_code[k] == 0 and np.abs(out_r[k] - o_r[k]).max() < 1e-4]
    assert len(both) >= 3

    s1 = fly.System(data, fly.Box(tuple(ex)), ids)
    nl = fly.List(s1, data.rcut())

    def gpu_point(r):
        nl.rebuild(r)
        return pot.energy(nl)[0], pot.hessian_eigenvalues(nl, mass_weight=True)  # Hessian and eigensolve on the device

    def cpu_point(r):
        osys.rebuild(otab.rcut, r, 8)
        return osys.energy(ids, 8), fly.hessian_eigenvalues(osys.hessian(ids, fz, 8, mass_weight=True))

    e_min, ev_min = gpu_point(r0)
    eo_min, evo_min = cpu_point(r0)
    n = 0
    for k in both[:4]:
        e_sp, ev_sp = gpu_point(out_r[k])
        eo_sp, evo_sp = cpu_point(o_r[k])
        assert abs((e_sp - e_min) - (eo_sp - eo_min)) < 1e-6
        nu, nu_o = fly.vineyard_prefactor(ev_min, ev_sp), fly.vineyard_prefactor(evo_min, evo_sp)
        assert (nu is None) == (nu_o is None)
        if nu is not None:
            assert abs(nu / nu_o - 1) < 1e-4
            assert 1e11 < nu < 1e16  # an attempt frequency, in Hz
            n += 1
    assert n >= 1
    s1.close()


def test_dimer_64_searches_branch_outs_classified(gpu, otab, oracle_lib, capsys):
    """64 searches on the apps/hydrogen-sized cell against the oracle. Trajectories are chaotic: a
    rounding-level difference can send a long search to another saddle. So every search is CLASSIFIED -- same exit and
    same saddle (then the barrier must agree to 1e-6 eV), or branched out -- and branch-outs must stay rare."""
    fly, ctx, data = gpu
    base, ex, ids = bcc_vacancy(6, a=2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    osys = oracle_lib.System(otab, ex)
    r0, _, failed, _ = osys.minimise(base, ids, fz, threads=8)
    assert not failed
    R = 64
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(R)]
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    sR = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
    out_r, out_ax, code, cnt = fly.Dimer().find_sp(sR, P, A, r0, fly.Generic(data))
    sR.close()
    o_r, o_ax, o_code, o_cnt = osys.find_sp_batch(P, A, ids, fz, r0, threads=8)
    E0 = _energy(osys, otab, ids, r0)
    same, branched = 0, []
    for k in range(R):
        if code[k] != o_code[k]:
            branched.append((k, int(code[k]), int(o_code[k]), int(cnt[k, 0]), int(o_cnt[k, 0])))
        elif code[k] == 0 and np.abs(out_r[k] - o_r[k]).max() > 1e-4:
            branched.append((k, 0, 0, int(cnt[k, 0]), int(o_cnt[k, 0])))
        else:
            same += 1
            if code[k] == 0:
                assert abs((_energy(osys, otab, ids, out_r[k]) - E0) - (_energy(osys, otab, ids, o_r[k]) - E0)) < 1e-6
    with capsys.disabled():
        print(f"\n[dimer parity] {same}/{R} searches identical in exit and saddle; branch-outs (replica, exit gpu, exit cpu, evals gpu, evals cpu): {branched}")
    assert len(branched) <= R // 10
    # a branch-out only happens on long trajectories: the short ones must all agree
    assert all(min(b[3], b[4]) > 150 for b in branched)


def test_queue_refill_equals_batch_equals_single(gpu, small_cell, oracle_lib):
    """ofly_dimer_find_sp_stream: 24 searches through 5 resident slots (finished slots re-arm on the device), the same 24
    as one batch of 24, as three shards of 8 (what three GPUs would run) and one by one: bit-identical records. What a
    search computes depends neither on the batch composition nor on the slot it ran in."""
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P8, A8 = small_cell
    starts = [osys.perturb(r0, fz, len(ids) - 1, (5, 3, 3, 4), rep) for rep in range(24)]
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    pot = fly.Generic(data)
    box = fly.Box(tuple(ex))

    def run(R, p, a):
        s = fly.System(data, box, ids, fz, replicas=R)
        try:
            return fly.Dimer().find_sp(s, p, a, r0, pot)
        finally:
            s.close()

    full = run(24, P, A)
    stream = run(5, P, A)
    for x, y in zip(full, stream):
        assert np.array_equal(x, y)
    from openfly_b200 import shard
    parts = {}
    for rank in range(3):
        mine, (p, a) = shard.shard((P, A), 24, 3, rank)
        o = run(len(mine), p, a)
        for k, g in enumerate(mine):
            parts[int(g)] = tuple(v[k] for v in o)
    for g in range(24):
        for x, y in zip(parts[g], (v[g] for v in full)):
            assert np.array_equal(x, y)
    for g in (0, 7, 23):
        one = run(1, P[g:g + 1], A[g:g + 1])
        for x, y in zip(one, full):
            assert np.array_equal(x[0], y[g])
    assert len(set(full[2].tolist())) >= 2  # the set is not trivial: several exit codes
    # fewer searches than slots
    few = run(16, P[:3], A[:3])
    for x, y in zip(few, full):
        assert np.array_equal(x, y[:3])
    # minimiser: same contract
    jit = np.stack([oracle_lib.jitter(base, (9, 2, 3, 40 + r), 0.05) for r in range(7)])
    s7 = fly.System(data, box, ids, fz, replicas=7)
    s2 = fly.System(data, box, ids, fz, replicas=2)
    a7, b7 = fly.LBFGS().minimise(s7, jit, pot), fly.LBFGS().minimise(s2, jit, pot)
    for x, y in zip(a7, b7):
        assert np.array_equal(x, y)
    s7.close(), s2.close()


def test_search_graph_is_reused_between_calls(gpu, small_cell):
    """The instantiated tick graph is kept while the kernel arguments are unchanged: a second identical call gives
    identical results with and without the cache."""
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    pot = fly.Generic(data)
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=len(P))
    a = fly.Dimer().find_sp(s, P, A, r0, pot)
    b = fly.Dimer().find_sp(s, P, A, r0, pot)
    os.environ["OFLY_NO_GRAPH_CACHE"] = "1"
    try:
        c = fly.Dimer().find_sp(s, P, A, r0, pot)
    finally:
        del os.environ["OFLY_NO_GRAPH_CACHE"]
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x, y) and np.array_equal(x, z)
    s.close()


def test_search_on_tile_path_equals_row_path(gpu, small_cell):
    """A whole batch of dimer searches with the cell-tiled kernels and with the row kernels (OFLY_TILES=0): same bits."""
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P, A = small_cell
    pot = fly.Generic(data)
    res = {}
    for mode in ("1", "0"):
        os.environ["OFLY_TILES"] = mode
        try:
            s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=len(P))
            res[mode] = fly.Dimer().find_sp(s, P, A, r0, pot)
            s.close()
        finally:
            del os.environ["OFLY_TILES"]
    for x, y in zip(res["1"], res["0"]):
        assert np.array_equal(x, y)


def test_staged_device_loop_equals_single_graph(gpu, small_cell, oracle_lib):
    """The device-side loop runs in stages with shrinking launch grids (64 -> 8 -> 1 bound on the active replicas); OFLY_STAGES=0
    keeps one graph for the whole call: bit-identical records."""
    fly, ctx, data = gpu
    base, ex, ids, fz, osys, r0, P8, A8 = small_cell
    starts = [osys.perturb(r0, fz, len(ids) - 1, (8, 3, 3, 4), rep) for rep in range(80)]
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    pot = fly.Generic(data)
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz, replicas=64)  # queue of 80 through 64 slots: refill, then three stages
    staged = fly.Dimer().find_sp(s, P, A, r0, pot)
    os.environ["OFLY_STAGES"] = "0"
    try:
        single = fly.Dimer().find_sp(s, P, A, r0, pot)
    finally:
        del os.environ["OFLY_STAGES"]
    again = fly.Dimer().find_sp(s, P, A, r0, pot)
    s.close()
    for x, y, z in zip(staged, single, again):
        assert np.array_equal(x, y) and np.array_equal(x, z)


def test_dimer_on_a_two_cta_team_matches_oracle(gpu, otab, oracle_lib):
    """A 1 458-atom cell (9^3): its 4 374-element vectors run on a cluster of two CTAs (search_api.cu pick_team), the one team
    shape between the single-CTA cells above and the 16-CTA teams of the C2 batch."""
    fly, ctx, data = gpu
    base, ex, ids = bcc_vacancy(9, a=2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    osys = oracle_lib.System(otab, ex)
    r0, _, failed, _ = osys.minimise(base, ids, fz, threads=8)
    assert not failed
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(4)]
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    code, cnt, o_cnt, n_ok = _check_dimer_against_oracle(fly, data, otab, ids, fz, ex, osys, r0, P, A, threads=4)
    assert n_ok >= 1
