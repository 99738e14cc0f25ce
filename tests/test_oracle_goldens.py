"""Pins the CPU oracle (oracle/fly_oracle.hpp) against every known-answer test the reference holds
for this path. Runs without a GPU."""
import numpy as np
import pytest

from tests.common import WEN, bcc_vacancy, golden_cell


@pytest.fixture(scope="module")
def O(oracle_lib):
    return oracle_lib


@pytest.fixture(scope="module")
def tab(O):
    return O.Tables(WEN)


def test_parser_goldens(O, tab):
    """tests/potential/EAM/data.cpp:28-47"""
    assert (tab.n_types, tab.numP, tab.numR) == (2, 10000, 10000)
    assert tab.rcut == 5.3
    assert tab.eval(O.KIND_F, 0, 0, 0, 0.0) == -0.1414216239132407
    assert tab.eval(O.KIND_F, 0, 0, 0, 0.02) == -0.2000010747904311
    assert np.allclose(tab.mass, [55.847, 1.008])


def test_eam_compute_goldens(O, tab):
    """tests/potential/EAM/eam.cpp:99-146 (energy uses the reference's `near`: atol 1e-10 | rtol 1e-4)."""
    pos, ex, ids = golden_cell()
    s = O.System(tab, ex)
    s.rebuild(tab.rcut, pos, 4)
    E = s.energy(ids, 4)
    assert abs(E - (-2748.5339306065985)) < 1e-9  # far tighter than the reference's own check
    g = s.gradient(ids, np.zeros(len(ids), np.uint8), 4)
    assert np.linalg.norm(g[0] - (-0.19182439085483005)) < 1e-10
    H = s.hessian(ids, np.zeros(len(ids), np.uint8), 4)
    ngd = np.array([[+9.0654715551510830, -1.0875103661087880, -1.0875103661087877],
                    [-1.0875103661087880, +9.0654715551510830, -1.0875103661087873],
                    [-1.0875103661087877, -1.0875103661087873, +9.0654715551510830]])
    assert np.sqrt((np.tril(H[:3, :3] - ngd) ** 2).sum()) < 1e-10
    # single-threaded == multi-threaded (eam.cpp:75-79 branches)
    assert np.array_equal(g, s.gradient(ids, np.zeros(len(ids), np.uint8), 1))


def test_eam_hess_goldens(O, tab):
    """tests/potential/EAM/eam.cpp:149-188"""
    pos, ex, ids = golden_cell(erase=False)
    s = O.System(tab, ex)
    s.rebuild(tab.rcut, pos, 4)
    g = s.gradient(ids, np.zeros(len(ids), np.uint8), 4)
    assert np.linalg.norm(g) < 1e-7
    H = s.hessian(ids, np.zeros(len(ids), np.uint8), 4)
    ev = np.linalg.eigvalsh(np.tril(H) + np.tril(H, -1).T)
    near0 = np.abs(ev) < 1e-10 + 1e-4 * 0  # `near(ev, 0.0)`
    assert near0[:3].all() and not near0[3:].any()


def test_adjacent_cells_golden(O):
    """tests/neigh/adjacent.cpp:23-94"""
    c = (1) + (1) * 5 + (1) * 25
    nb = O.adjacent([5, 5, 5], c)
    expect = [x + y * 5 + z * 25 for z in (0, 1, 2) for y in (0, 1, 2) for x in (0, 1, 2) if (x, y, z) != (1, 1, 1)]
    assert nb.tolist() == expect
    assert len(O.adjacent([5, 5, 5], 2 + 2 * 5)) == 17
    assert len(O.adjacent([5, 5, 5], 2)) == 11
    assert len(O.adjacent([5, 5, 5], 0)) == 7


def test_hypergrid_golden(O):
    """tests/system/boxes/hypergrid.cpp:24-44"""
    shape, cell = O.grid_info([10, 10, 10], 3)
    assert shape.tolist() == [5, 5, 5]
    idx = O.cell_idx([10, 10, 10], 3, [[0, 0, 0], [5, 5, 5], [9.5, 9.5, 9.5]])
    assert idx.tolist() == [1 + 5 + 25, 2 + 10 + 50, 3 + 15 + 75]


def test_gen_image_golden(O):
    """tests/system/boxes/orthorhombic.cpp:52-99"""
    ex = [10, 10, 10]
    assert O.gen_image(ex, 3, [5, 5, 5], 0, +1) is None
    origin = np.array([0.2, 0.0, 0.0])
    assert O.gen_image(ex, 3, origin, 0, -1) is None
    im = O.gen_image(ex, 3, origin, 0, +1)
    assert im is not None and np.linalg.norm(im - (origin + [10, 0, 0])) < 1e-3
    origin = np.array([9.8, 0.0, 0.0])
    assert O.gen_image(ex, 3, origin, 0, +1) is None
    im = O.gen_image(ex, 3, origin, 0, -1)
    assert im is not None and np.linalg.norm(im - (origin - [10, 0, 0])) < 1e-3


def test_canon_image_fuzz(O):
    """tests/system/boxes/orthorhombic.cpp:24-50 (fully periodic slice)."""
    rng = np.random.default_rng(33)
    for _ in range(200):
        ex = rng.random(3) + 1
        a = rng.random(3) * ex
        b = rng.random(3) * ex
        bp = b + np.floor(10 * rng.random(3)) * ex
        c = O.canon_image(ex, bp)[0]
        assert np.all(c >= 0) and np.all(c < ex)
        assert abs(np.linalg.norm(c - a) - np.linalg.norm(a - b)) < 1e-3


def test_spline_goldens(O):
    """tests/utility/spline.cpp:27-83"""
    y = np.ones(100)
    x = np.linspace(0.0, 0.98, 57)
    assert np.allclose(O.spline_eval(y, 0.01, 0, x), 1.0, atol=1e-12)
    assert np.allclose(O.spline_eval(y, 0.01, 1, x), 0.0, atol=1e-10)
    assert np.allclose(O.spline_eval(y, 0.01, 2, x), 0.0, atol=1e-8)
    rng = np.random.default_rng(5)
    a, b, c, d = rng.random(4)
    f = lambda t: a * t ** 3 + b * t ** 2 + c * t + d  # noqa: E731
    n = 1000
    xs = np.linspace(0, 1, n + 1)
    q = np.linspace(0.1, 0.9, 33)
    assert np.allclose(O.spline_eval(f(xs), 1.0 / n, 0, q), f(q), rtol=1e-2)
    assert np.allclose(O.spline_eval(f(xs), 1.0 / n, 1, q), 3 * a * q ** 2 + 2 * b * q + c, rtol=1e-2)
    with pytest.raises(O.OutOfBounds):
        O.spline_eval(y, 0.01, 0, [1.05])  # beyond the 5 padded knots (spline.hpp:141-143)


def _brute_force(pos, ex, periodic, r_cut):
    """tests/neigh/list.cpp:42-97: O(N^2 * 27) minimum-image list (index + image vector)."""
    shifts = []
    for sz in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sx in (-1, 0, 1):
                s = np.array([sx, sy, sz], dtype=float) * np.asarray(periodic, dtype=float)
                off = s * ex
                if (sx, sy, sz) == (0, 0, 0) or np.any(off != 0):
                    shifts.append(off)
    shifts = np.unique(np.array(shifts), axis=0)
    out = []
    for i in range(len(pos)):
        d = (pos[None, :, :] + shifts[:, None, :]) - pos[i]      # [S, N, 3]
        r = np.sqrt((d * d).sum(-1))
        best = r.argmin(0)
        rmin = r[best, np.arange(len(pos))]
        js = [j for j in range(len(pos)) if j != i and rmin[j] ** 2 < r_cut * r_cut]
        out.append([(j, d[best[j], j]) for j in js])
    return out


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_list_fuzz_vs_brute_force(O, tab, seed):
    """tests/neigh/list.cpp:167-208 restricted to orthorhombic boxes (the in-scope Box alternative)."""
    rng = np.random.default_rng(1334 + seed)
    ex = 1 + 9 * rng.random(3) + rng.random(3)
    periodic = tuple(bool(x) for x in rng.random(3) < 0.5)
    n = int(rng.integers(100, 300))
    pos = rng.random((n, 3)) * ex
    r_cut = ex.min() / 3
    s = O.System(tab, ex, periodic)
    for threads in (1, 4):
        s.rebuild(r_cut, pos, threads)
        ref = _brute_force(pos, ex, periodic, r_cut)
        for i in range(n):
            idx, r, dr = s.for_neighbours(i, r_cut)
            order = np.argsort(idx, kind="stable")
            assert idx[order].tolist() == [j for j, _ in ref[i]]
            for k, (j, d) in zip(order, ref[i]):
                assert np.linalg.norm(dr[k] - d) < 1e-10


def test_list_update_matches_rebuild(O, tab):
    """tests/neigh/list.cpp:210-251: after update() of sub-skin displacements the (filtered) list is still exact."""
    rng = np.random.default_rng(1111)
    pos, ex, ids = bcc_vacancy(4, jitter=0.05)
    r_cut, skin = 3.0, 0.4
    s = O.System(tab, ex)
    s.rebuild(r_cut + skin, pos)
    delta = rng.uniform(-1, 1, pos.shape) * (skin / 2 / np.sqrt(3)) * 0.99
    s.update(delta)
    s2 = O.System(tab, ex)
    s2.rebuild(r_cut, pos - delta)
    for i in range(0, len(pos), 7):
        a = sorted(s.for_neighbours(i, r_cut)[0].tolist())
        b = sorted(s2.for_neighbours(i, r_cut)[0].tolist())
        assert a == b


def test_lbfgs_and_dimer_smoke(O, tab):
    """No reference tests exist for minimise/ and saddle/ (SURVEY.md 0.5): sanity properties only."""
    pos, ex, ids = bcc_vacancy(3, a=2.855, remove=(1,), hydrogen=False, jitter=0.02)
    fz = np.zeros(len(ids), np.uint8)
    s = O.System(tab, ex)
    r, g, failed, cnt = s.minimise(pos, ids, fz)
    assert not failed and np.linalg.norm(g) < 1e-6
    s.rebuild(tab.rcut, pos)
    e0 = s.energy(ids)
    s.rebuild(tab.rcut, r)
    assert s.energy(ids) < e0
    p, ax = s.perturb(r, fz, 0, (1, 3, 3, 4), 0)
    assert abs(np.linalg.norm(ax) - 1) < 1e-12
    out_r, out_ax, code, cnt = s.find_sp(p, ax, ids, fz, r)
    assert code in (0, -1, -2, -3)
    assert abs(np.linalg.norm(out_ax) - 1) < 1e-6
