"""GPU parity: neigh::List and potential::EAM through the C ABI vs the CPU oracle and the reference goldens."""
import numpy as np
import pytest

from tests.common import WEN, bcc_vacancy, golden_cell, pair_multiset

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


def test_tables_bit_exact(gpu, otab, oracle_lib):
    fly, ctx, data = gpu
    O = oracle_lib
    assert (data.n_types, data.knots_rho, data.knots_r) == (2, 10004, 10004)
    assert data.rcut() == otab.rcut
    for kind, i, j in [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1), (2, 0, 0), (2, 1, 0), (2, 0, 1), (2, 1, 1)]:
        assert np.array_equal(data.knots(kind, i, j), otab.knots(kind, i, j)), (kind, i, j)


def _compare_lists(fly, nl, osys, r_list, pos, replica=0):
    osys.rebuild(r_list, pos)
    c0, i0, s0 = osys.export()
    c1, i1, s1 = nl.export(replica)
    assert np.array_equal(c0, c1)
    assert pair_multiset(c0, i0, s0) == pair_multiset(c1, i1, s1)  # bit-exact neighbour sets
    assert np.array_equal(osys.cells(), nl.cells(replica))          # bit-exact cell ids


def test_golden_energy_gradient_hessian(gpu, otab, oracle_lib):
    """Known answers of tests/potential/EAM/eam.cpp:99-146."""
    fly, ctx, data = gpu
    pos, ex, ids = golden_cell()
    s = fly.System(data, fly.Box(tuple(ex)), ids)
    nl = fly.List(s, data.rcut())
    nl.rebuild(pos)
    pot = fly.Generic(data)
    E = pot.energy(nl)[0]
    assert abs(E - (-2748.5339306065985)) <= 1e-10 * abs(E)
    g = pot.gradient(nl)[0]
    assert np.linalg.norm(g[0] - (-0.19182439085483005)) < 1e-10
    H = pot.hessian(nl)
    ngd = np.array([[+9.0654715551510830, -1.0875103661087880, -1.0875103661087877],
                    [-1.0875103661087880, +9.0654715551510830, -1.0875103661087873],
                    [-1.0875103661087877, -1.0875103661087873, +9.0654715551510830]])
    assert np.abs(np.tril(H[:3, :3] - ngd)).max() < 1e-10
    # whole Hessian against the oracle (lower triangle)
    osys = oracle_lib.System(otab, ex)
    osys.rebuild(otab.rcut, pos)
    Ho = osys.hessian(ids, np.zeros(len(ids), np.uint8), threads=8)
    assert np.abs(np.tril(H) - np.tril(Ho)).max() < 1e-9
    _compare_lists(fly, nl, osys, otab.rcut, pos)
    s.close()


def test_perfect_cell_zero_modes(gpu):
    """tests/potential/EAM/eam.cpp:149-188: |g| < 1e-7 and exactly 3 zero Hessian modes."""
    fly, ctx, data = gpu
    pos, ex, ids = golden_cell(erase=False)
    s = fly.System(data, fly.Box(tuple(ex)), ids)
    nl = fly.List(s, data.rcut())
    nl.rebuild(pos)
    pot = fly.Generic(data)
    assert np.linalg.norm(pot.gradient(nl)[0]) < 1e-7
    H = pot.hessian(nl)
    ev = np.linalg.eigvalsh(np.tril(H) + np.tril(H, -1).T)
    assert np.all(np.abs(ev[:3]) < 1e-8) and np.all(np.abs(ev[3:]) > 1e-4)
    s.close()


@pytest.mark.parametrize("n,hyd,r_extra", [(10, False, 0.0), (6, True, 0.1731), (4, True, 0.0)])
def test_parity_vs_oracle(gpu, otab, oracle_lib, n, hyd, r_extra):
    """C1-style (2000-atom) and C3-style (small box, many ghosts, Fe + H) cells: sets, cells, E, g, update."""
    fly, ctx, data = gpu
    pos, ex, ids = bcc_vacancy(n, remove=(1,), jitter=0.05, hydrogen=hyd)
    fz = np.zeros(len(ids), np.uint8)
    fz[5] = 1  # one frozen atom: zero gradient but still a density source
    r_list = data.rcut() + r_extra
    s = fly.System(data, fly.Box(tuple(ex)), ids, fz)
    nl = fly.List(s, r_list)
    nl.rebuild(pos)
    osys = oracle_lib.System(otab, ex)
    _compare_lists(fly, nl, osys, r_list, pos)
    pot = fly.Generic(data)
    E, Eo = pot.energy(nl)[0], osys.energy(ids, 4)
    assert abs(E - Eo) <= 1e-10 * abs(Eo)
    g, go = pot.gradient(nl)[0], osys.gradient(ids, fz, 4)
    assert np.abs(g - go).max() < 1e-9
    assert np.all(g[5] == 0)
    # update(): r <- r - delta on both sides, same lists
    rng = np.random.default_rng(7)
    delta = rng.normal(0, 0.01, pos.shape)
    nl.update(delta)
    osys.update(delta)
    g, go = pot.gradient(nl)[0], osys.gradient(ids, fz, 4)
    assert np.abs(g - go).max() < 1e-9
    E, Eo = pot.energy(nl)[0], osys.energy(ids, 4)
    assert abs(E - Eo) <= 1e-10 * abs(Eo)
    s.close()


def test_batched_replicas_independent(gpu, otab, oracle_lib):
    fly, ctx, data = gpu
    base, ex, ids = bcc_vacancy(6, a=2.855, remove=(1,), hydrogen=True)
    R = 5
    pos = np.stack([oracle_lib.jitter(base, (1, 2, 3, 4 + r), 0.04) for r in range(R)])
    s = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
    nl = fly.List(s, data.rcut())
    nl.rebuild(pos)
    pot = fly.Generic(data)
    E = pot.energy(nl)
    g = pot.gradient(nl)
    osys = oracle_lib.System(otab, ex)
    for r in range(R):
        _compare_lists(fly, nl, osys, otab.rcut, pos[r], replica=r)
        Eo = osys.energy(ids, 2)
        assert abs(E[r] - Eo) <= 1e-10 * abs(Eo)
        assert np.abs(g[r] - osys.gradient(ids, np.zeros(len(ids), np.uint8), 2)).max() < 1e-9
    s.close()


def test_partially_periodic_and_fuzz_boxes(gpu, otab, oracle_lib):
    """Orthorhombic slice of tests/neigh/list.cpp:167-208: random boxes, random periodicity, r = min width / 3."""
    fly, ctx, data = gpu
    rng = np.random.default_rng(1334)
    for trial in range(12):
        ex = 1 + 9 * rng.random(3) + rng.random(3)
        periodic = tuple(bool(x) for x in rng.random(3) < 0.5)
        n = int(rng.integers(100, 1000))
        pos = rng.random((n, 3)) * ex
        r_cut = ex.min() / 3
        ids = np.zeros(n, np.uint32)
        s = fly.System(data, fly.Box(tuple(ex), periodic), ids)
        nl = fly.List(s, r_cut)
        nl.rebuild(pos)
        osys = oracle_lib.System(otab, ex, periodic)
        _compare_lists(fly, nl, osys, r_cut, pos)
        s.close()


def test_errors(gpu, oracle_lib):
    fly, ctx, data = gpu
    pos, ex, ids = bcc_vacancy(4, hydrogen=False)
    s = fly.System(data, fly.Box(tuple(ex)), ids)
    pot = fly.Generic(data)
    nl = fly.List(s, 3.0)  # list radius below the potential cut-off: for_neighbours throws (list.hpp:141-143)
    nl.rebuild(pos)
    with pytest.raises(fly.OflyError):
        pot.gradient(nl)
    with pytest.raises(fly.OflyError):
        fly.List(s, 100.0).rebuild(pos)  # r_cut too big for the box (hypergrid.hpp:46)
    # two atoms on top of each other: rho overruns the F table -> Spline::OutOfBounds
    bad = pos.copy()
    bad[1:40] = bad[0] + np.linspace(0.01, 0.4, 39)[:, None]
    nl = fly.List(s, data.rcut())
    nl.rebuild(bad)
    with pytest.raises(fly.OutOfBounds):
        pot.gradient(nl)
    with pytest.raises(fly.OflyError):
        fly.System(data, fly.Box(tuple(ex)), np.full(10, 7, np.uint32))  # TypeID outside the table
    s.close()


def test_hessian_eigenvalues_on_device(gpu, otab, oracle_lib):
    """Hessian::eigenvalues() (hessian.hpp:86-93) solved on the device (cuSOLVER) against numpy on the oracle's Hessian:
    3 zero modes of the vacancy-free cell (tests/potential/EAM/eam.cpp:173-183), the rest to 1e-9 absolute."""
    fly, ctx, data = gpu
    pos, ex, ids = golden_cell(erase=False)
    s = fly.System(data, fly.Box(tuple(ex)), ids)
    nl = fly.List(s, data.rcut())
    nl.rebuild(pos)
    pot = fly.Generic(data)
    ev = pot.hessian_eigenvalues(nl, mass_weight=False)
    assert np.all(np.diff(ev) > -1e-12)  # ascending (degenerate modes may swap at rounding level)
    assert np.abs(ev[:3]).max() < 1e-10 and ev[3] > 1e-4
    osys = oracle_lib.System(otab, ex)
    osys.rebuild(otab.rcut, pos, 8)
    Ho = osys.hessian(ids, np.zeros(len(ids), np.uint8), 8)
    assert np.abs(ev - fly.hessian_eigenvalues(Ho)).max() < 1e-9
    evm = pot.hessian_eigenvalues(nl, mass_weight=True)
    Hm = osys.hessian(ids, np.zeros(len(ids), np.uint8), 8, mass_weight=True)
    assert np.abs(evm - fly.hessian_eigenvalues(Hm)).max() < 1e-10
    s.close()


def test_sums_do_not_depend_on_subwarp_width(gpu):
    """Every per-atom sum is defined over 16 virtual lanes (eam.cuh, VirtSum): 4, 8 or 16 physical lanes per atom give
    the same bits, so energies / forces do not depend on the batch size (which picks the width)."""
    import os

    fly, ctx, data = gpu
    pos, ex, ids = bcc_vacancy(6, a=2.855, remove=(1,), jitter=0.05, hydrogen=True)
    s = fly.System(data, fly.Box(tuple(ex)), ids, replicas=3)
    nl = fly.List(s, data.rcut() + 0.17)
    nl.rebuild(np.stack([pos, pos + 0.01, pos - 0.02]))
    pot = fly.Generic(data)
    got = {}
    try:
        for G in (4, 8, 16):
            os.environ["OFLY_GROUP"] = str(G)
            got[G] = (pot.energy(nl), pot.gradient(nl))
    finally:
        del os.environ["OFLY_GROUP"]
    for G in (8, 16):
        assert np.array_equal(got[G][0], got[4][0]) and np.array_equal(got[G][1], got[4][1])
    s.close()
    # and a replica alone gives the bits it gives inside a batch of 3 (other width, other grid)
    s1 = fly.System(data, fly.Box(tuple(ex)), ids)
    nl1 = fly.List(s1, data.rcut() + 0.17)
    nl1.rebuild(pos + 0.01)
    assert np.array_equal(pot.gradient(nl1)[0], got[4][1][1]) and pot.energy(nl1)[0] == got[4][0][1]
    s1.close()


def test_rows_are_ordered_by_neighbour_index(gpu):
    """Step 6 of the rebuild (neigh.cuh, k_nb_order_rows): every stored row is ordered by (neighbour index, image shift)
    whatever slot order the atomics of the cell binning produced, so two rebuilds from the same positions give the same
    rows, and gradients computed on them the same bits; rows of every length class (<= 32, <= 64, <= 128 entries sorted in
    registers, longer ones in shared memory) take it."""
    fly, ctx, data = gpu
    pot = fly.Generic(data)
    for n, extra, lo, hi in ((5, 0.17, 33, 64), (6, -1.2, 1, 32), (6, 1.2, 65, 128), (6, 2.0, 129, 512)):
        pos, ex, ids = bcc_vacancy(n, a=2.855, remove=(1,), jitter=0.05, hydrogen=True)
        s = fly.System(data, fly.Box(tuple(ex)), ids, replicas=2)
        nl = fly.List(s, data.rcut() + extra)
        P = np.stack([pos, pos + 0.01])
        seen = []
        for _ in range(2):
            nl.rebuild(P)
            cnt, idx, sh = nl.export(1)
            off = np.concatenate([[0], np.cumsum(cnt)])
            key = idx.astype(np.int64) * 32 + sh
            for i in range(len(cnt)):
                row = key[off[i]:off[i + 1]]
                assert (np.diff(row) > 0).all(), f"row {i} is not strictly ordered"
            seen.append((cnt, idx, sh, pot.gradient(nl) if extra > 0 else None))
        assert all(np.array_equal(x, y) for x, y in zip(seen[0][:3], seen[1][:3]))
        if extra > 0:
            assert np.array_equal(seen[0][3], seen[1][3])
        assert lo <= np.median(cnt) <= hi, (np.median(cnt), lo, hi)
        s.close()


def test_tables_from_memory_and_from_knots(gpu):
    """DataEAM(std::istream) (data.cpp:67) and tables handed over as knots: same bits as the path constructor; EAM's
    TypeMap check (eam.hpp:50-63)."""
    fly, ctx, data = gpu
    mem = fly.DataEAM(ctx, text=open(WEN, "rb").read())
    sel = [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1), (2, 0, 0), (2, 1, 0), (2, 1, 1)]
    for kind, i, j in sel:
        assert np.array_equal(mem.knots(kind, i, j), data.knots(kind, i, j))
    assert mem.names == ["Fe", "H"] and mem.rcut() == data.rcut()
    n = data.n_types
    F = np.stack([data.knots(0, i) for i in range(n)])
    phi = np.zeros((n * n, data.knots_r, 4))
    V = np.zeros((n * n, data.knots_r, 4))
    for i in range(n):
        for j in range(n):
            phi[i + n * j] = data.knots(1, i, j)
            V[max(i, j) + n * min(i, j)] = data.knots(2, i, j)
    kn = fly.DataEAM.from_knots(ctx, data.names, data.mass, data.d_rho, data.d_r, data.rcut(), F, phi, V)
    pos, ex, ids = bcc_vacancy(4, jitter=0.03, hydrogen=True)
    out = []
    for d in (data, mem, kn):
        s = fly.System(d, fly.Box(tuple(ex)), ids)
        nl = fly.List(s, d.rcut())
        nl.rebuild(pos)
        pot = fly.Generic(d, type_map=["Fe", "H"])
        out.append((pot.energy(nl), pot.gradient(nl)))
        s.close()
    for e, g in out[1:]:
        assert np.array_equal(e, out[0][0]) and np.array_equal(g, out[0][1])
    with pytest.raises(fly.OflyError, match="Not enough types"):
        fly.Generic(data, type_map=["Fe"])
    with pytest.raises(fly.OflyError, match="TypeID mismatch"):
        fly.Generic(data, type_map=["H", "Fe"])
    fly.Generic(data, type_map=["Fe", "H", "He"])  # more types than the table is fine
    mem.close(), kn.close()


def test_out_of_table_negative_density(gpu):
    """Spline::fetch converts x / dx with static_cast<size_t> (spline.hpp:135-143): an argument <= -dx (a negative embedding
    density from a table whose phi goes negative) is an OutOfBounds in the reference, never a read before the table."""
    fly, ctx, data = gpu
    n = data.n_types
    F = np.stack([data.knots(0, i) for i in range(n)])
    phi = np.zeros((n * n, data.knots_r, 4))
    V = np.zeros((n * n, data.knots_r, 4))
    for i in range(n):
        for j in range(n):
            phi[i + n * j] = -data.knots(1, i, j)  # every density contribution negated
            V[max(i, j) + n * min(i, j)] = data.knots(2, i, j)
    neg = fly.DataEAM.from_knots(ctx, data.names, data.mass, data.d_rho, data.d_r, data.rcut(), F, phi, V)
    pos, ex, ids = bcc_vacancy(4, hydrogen=False)
    s = fly.System(neg, fly.Box(tuple(ex)), ids)
    nl = fly.List(s, neg.rcut())
    nl.rebuild(pos)
    pot = fly.Generic(neg)
    with pytest.raises(fly.OutOfBounds):
        pot.gradient(nl)
    with pytest.raises(fly.OutOfBounds):
        pot.energy(nl)
    s.close()
    s = fly.System(data, fly.Box(tuple(ex)), ids)  # the context survived
    nl = fly.List(s, data.rcut())
    nl.rebuild(pos)
    assert np.isfinite(fly.Generic(data).energy(nl)[0])
    s.close(), neg.close()


def test_tile_path_equals_row_path(gpu):
    """The cell-tiled kernels (tile.cuh: window staged in shared memory, 16-bit slot rows) against the row kernels
    (OFLY_TILES=0): identical neighbour rows, and identical bits for energy and gradient -- on the C3-sized cell (3 real cells
    per axis: partial tiles, windows that span almost the whole grid), a C1-sized cell and an odd, non-cubic box."""
    import os

    fly, ctx, data = gpu
    cases = [bcc_vacancy(6, a=2.855, remove=(1,), jitter=0.05, hydrogen=True), bcc_vacancy(10, jitter=0.05, hydrogen=False)]
    rng = np.random.default_rng(11)
    ex = np.array([23.0, 17.5, 29.1])
    cases.append((rng.random((1500, 3)) * ex * 0.999 + 0.5 * (rng.random((1500, 3)) - 0.5) * 0, ex, np.zeros(1500, np.uint32)))
    for pos, ex, ids in cases:
        if len(ids) == 1500:  # random gas: thin it so that no two atoms overlap (keeps rho inside the table)
            keep = [0]
            for k in range(1, len(pos)):
                d = pos[keep] - pos[k]
                d -= ex * np.round(d / ex)
                if (d * d).sum(1).min() > 2.0 ** 2:
                    keep.append(k)
            pos, ids = pos[keep], ids[: len(keep)]
        R = 2
        P = np.stack([pos, pos + 0.02 * np.cos(np.arange(pos.size).reshape(pos.shape))])
        out = {}
        for mode in ("1", "0"):
            os.environ["OFLY_TILES"] = mode
            try:
                s = fly.System(data, fly.Box(tuple(ex)), ids, replicas=R)
                nl = fly.List(s, data.rcut() + 0.17)
                nl.rebuild(P)
                pot = fly.Generic(data)
                d = 0.01 * np.sin(np.arange(P.size).reshape(P.shape))
                g0 = pot.gradient(nl)
                nl.update(d)
                out[mode] = (nl.export(0), nl.export(1), pot.energy(nl), g0, pot.gradient(nl))
                s.close()
            finally:
                del os.environ["OFLY_TILES"]
        a, b = out["1"], out["0"]
        for k in (0, 1):
            for x, y in zip(a[k], b[k]):
                assert np.array_equal(x, y)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
        assert np.abs(a[3]).max() > 1e-3
