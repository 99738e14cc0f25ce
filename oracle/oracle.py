# SPDX-License-Identifier: GPL-3.0-or-later
"""TEST INFRASTRUCTURE ONLY -- ctypes view of the CPU oracle (oracle/liboracle*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (openfly_b200/) must never import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_fast.so with the committed Makefile."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE, "-j2"], check=True, capture_output=True)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleError(RuntimeError):
    pass


class OutOfBounds(OracleError):
    """Spline::OutOfBounds (include/libfly/utility/spline.hpp:46-51)."""


class Lib:
    def __init__(self, fast: bool = False, path: str | None = None):
        name = "liboracle_fast.so" if fast else "liboracle.so"
        path = path or os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        self.l = C.CDLL(path)
        l = self.l
        l.orc_last_error.restype = C.c_char_p
        l.orc_tables_load.restype = C.c_void_p
        l.orc_tables_load.argtypes = [C.c_char_p]
        l.orc_tables_free.argtypes = [C.c_void_p]
        l.orc_tables_info.argtypes = [C.c_void_p, _lp, _dp]
        l.orc_tables_num_knots.restype = C.c_long
        l.orc_tables_num_knots.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        l.orc_tables_knots.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp]
        l.orc_tables_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp]
        l.orc_spline_eval.argtypes = [_dp, C.c_long, C.c_double, C.c_int, _dp, C.c_long, _dp]
        l.orc_grid_info.argtypes = [_dp, C.c_double, _ip, _dp]
        l.orc_cell_idx.argtypes = [_dp, C.c_double, _dp, C.c_long, _ip]
        l.orc_canon_image.argtypes = [_dp, _dp, C.c_long, _dp]
        l.orc_gen_image.argtypes = [_dp, C.c_double, _dp, C.c_int, C.c_int, _dp]
        l.orc_adjacent.argtypes = [_ip, C.c_int, _ip]
        l.orc_system_create.restype = C.c_void_p
        l.orc_system_create.argtypes = [C.c_void_p, _dp, _ip]
        l.orc_system_free.argtypes = [C.c_void_p]
        l.orc_list_rebuild.argtypes = [C.c_void_p, C.c_double, _dp, C.c_long, C.c_int]
        l.orc_list_update.argtypes = [C.c_void_p, _dp, C.c_long]
        l.orc_list_num_plus_ghosts.restype = C.c_long
        l.orc_list_num_plus_ghosts.argtypes = [C.c_void_p]
        l.orc_list_export.restype = C.c_long
        l.orc_list_export.argtypes = [C.c_void_p, _ip, _ip, _ip, C.c_long]
        l.orc_list_cells.argtypes = [C.c_void_p, _ip]
        l.orc_list_for_neighbours.restype = C.c_long
        l.orc_list_for_neighbours.argtypes = [C.c_void_p, C.c_long, C.c_double, _ip, _dp, _dp, C.c_long]
        l.orc_eam_energy.argtypes = [C.c_void_p, _u32p, C.c_long, C.c_int, _dp]
        l.orc_eam_gradient.argtypes = [C.c_void_p, _u32p, _u8p, C.c_long, C.c_int, _dp]
        l.orc_eam_hessian.argtypes = [C.c_void_p, _u32p, _u8p, C.c_long, C.c_int, _dp, C.c_int]
        l.orc_lbfgs_minimise.argtypes = [C.c_void_p, _ip, _dp, _dp, _u32p, _u8p, C.c_long, C.c_int, _dp, _dp, _ip, _lp]
        l.orc_dimer_find_sp.argtypes = [C.c_void_p, _ip, _dp, _dp, _dp, _u32p, _u8p, C.c_long, _dp, _dp, C.c_long,
                                        C.c_double, C.c_int, C.c_int, _dp, _dp, _ip, _lp]
        l.orc_dimer_find_sp_batch.argtypes = [C.c_void_p, _ip, _dp, _dp, _dp, _u32p, _u8p, C.c_long, C.c_long, _dp, _dp,
                                              C.c_long, C.c_double, C.c_int, C.c_int, _dp, _dp, _ip, _lp]
        l.orc_lbfgs_minimise_batch.argtypes = [C.c_void_p, _ip, _dp, _dp, _u32p, _u8p, C.c_long, C.c_long, C.c_int, _dp, _dp,
                                               _ip, _lp]
        l.orc_bcc_lattice.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, _dp]
        l.orc_jitter.argtypes = [_dp, C.c_long, _u64p, C.c_double]
        l.orc_choose.argtypes = [C.c_long, C.c_long, _u64p, _lp]
        l.orc_perturb.argtypes = [C.c_void_p, _dp, _u8p, C.c_long, C.c_long, _u64p, C.c_long, C.c_double, C.c_double, _dp, _dp]
        l.orc_time_gradient.restype = C.c_double
        l.orc_time_gradient.argtypes = [C.c_void_p, _u32p, _u8p, C.c_long, _dp, C.c_int, C.c_int, _dp]
        l.orc_time_rebuild.restype = C.c_double
        l.orc_time_rebuild.argtypes = [C.c_void_p, C.c_double, _dp, C.c_long, C.c_int, C.c_int]

    def check(self, rc: int) -> None:
        if rc == 2:
            raise OutOfBounds(self.l.orc_last_error().decode())
        if rc != 0:
            raise OracleError(self.l.orc_last_error().decode())

    def max_threads(self) -> int:
        return int(self.l.orc_max_threads())


_libs: dict = {}


def lib(fast: bool = False, path: str | None = None) -> Lib:
    key = (fast, path)
    if key not in _libs:
        _libs[key] = Lib(fast, path)
    return _libs[key]


KIND_F, KIND_PHI, KIND_V = 0, 1, 2


class Tables:
    """DataEAM (src/potential/EAM/data.cpp:67-153)."""

    def __init__(self, path: str, fast: bool = False, libpath: str | None = None):
        self.L = lib(fast, libpath)
        self.h = self.L.l.orc_tables_load(path.encode())
        if not self.h:
            raise OracleError(self.L.l.orc_last_error().decode())
        info = np.zeros(3, dtype=np.int64)
        dinfo = np.zeros(3 + 16, dtype=np.float64)
        self.L.l.orc_tables_info(self.h, _ptr(info, _lp), _ptr(dinfo, _dp))
        self.n_types, self.numP, self.numR = (int(x) for x in info)
        self.delP, self.delR, self.rcut = (float(x) for x in dinfo[:3])
        self.mass = dinfo[3:3 + self.n_types].copy()

    def knots(self, kind: int, i: int, j: int = 0) -> np.ndarray:
        n = self.L.l.orc_tables_num_knots(self.h, kind, i, j)
        out = np.zeros((n, 4), dtype=np.float64)
        self.L.l.orc_tables_knots(self.h, kind, i, j, _ptr(out, _dp))
        return out

    def eval(self, kind: int, i: int, j: int, which: int, x: float) -> float:
        out = C.c_double()
        self.L.check(self.L.l.orc_tables_eval(self.h, kind, i, j, which, x, C.byref(out)))
        return out.value


def spline_eval(y, dx: float, which: int, x) -> np.ndarray:
    L = lib()
    y = _f64(y)
    x = _f64(np.atleast_1d(x))
    out = np.zeros_like(x)
    L.check(L.l.orc_spline_eval(_ptr(y, _dp), len(y), dx, which, _ptr(x, _dp), len(x), _ptr(out, _dp)))
    return out


def grid_info(extents, r_cut: float):
    L = lib()
    ex = _f64(extents)
    shape = np.zeros(3, dtype=np.int32)
    cell = np.zeros(3, dtype=np.float64)
    L.check(L.l.orc_grid_info(_ptr(ex, _dp), r_cut, _ptr(shape, _ip), _ptr(cell, _dp)))
    return shape, cell


def cell_idx(extents, r_cut: float, x) -> np.ndarray:
    L = lib()
    ex = _f64(extents)
    x = _f64(x).reshape(-1, 3)
    out = np.zeros(len(x), dtype=np.int32)
    L.check(L.l.orc_cell_idx(_ptr(ex, _dp), r_cut, _ptr(x, _dp), len(x), _ptr(out, _ip)))
    return out


def canon_image(extents, x) -> np.ndarray:
    L = lib()
    ex = _f64(extents)
    x = _f64(x).reshape(-1, 3)
    out = np.zeros_like(x)
    L.l.orc_canon_image(_ptr(ex, _dp), _ptr(x, _dp), len(x), _ptr(out, _dp))
    return out


def gen_image(extents, r_cut: float, x, ax: int, sign: int):
    L = lib()
    ex = _f64(extents)
    x = _f64(x)
    out = np.zeros(3)
    ok = L.l.orc_gen_image(_ptr(ex, _dp), r_cut, _ptr(x, _dp), ax, sign, _ptr(out, _dp))
    return out if ok else None


def adjacent(shape, cell: int) -> np.ndarray:
    L = lib()
    sh = np.ascontiguousarray(shape, dtype=np.int32)
    out = np.zeros(26, dtype=np.int32)
    n = L.l.orc_adjacent(_ptr(sh, _ip), cell, _ptr(out, _ip))
    return out[:n].copy()


def bcc_lattice(a: float, nx: int, ny: int, nz: int) -> np.ndarray:
    """motif_to_lattice for the bcc motif (include/libfly/utility/lattice.hpp:50-109)."""
    out = np.zeros((2 * nx * ny * nz, 3), dtype=np.float64)
    lib().l.orc_bcc_lattice(a, nx, ny, nz, _ptr(out, _dp))
    return out


def jitter(r: np.ndarray, seed, stddev: float) -> np.ndarray:
    r = _f64(r).copy()
    s = np.asarray(seed, dtype=np.uint64)
    lib().l.orc_jitter(_ptr(r, _dp), r.size, _ptr(s, _u64p), stddev)
    return r


def choose(n: int, k: int, seed) -> np.ndarray:
    s = np.asarray(seed, dtype=np.uint64)
    out = np.zeros(k, dtype=np.int64)
    lib().l.orc_choose(n, k, _ptr(s, _u64p), _ptr(out, _lp))
    return out


DIMER_IOPT = np.array([10, 1000, 5, 1, 6, 20, 1], dtype=np.int32)
DIMER_DOPT = np.array([1e-5, 1.1, 0.0, 0.5, 0.05, 1.5, 0.5, 3.0, 0.001, np.pi / 360.0], dtype=np.float64)
LBFGS_IOPT = np.array([10, 2000], dtype=np.int32)
LBFGS_DOPT = np.array([1e-6, 1.1, 0.0, 0.5, 0.01, 1.5, 0.5], dtype=np.float64)


class System:
    """Tables + orthorhombic box + one neigh::List; the reference's (pot, nl) pair."""

    def __init__(self, tables: Tables, extents, periodic=(True, True, True)):
        self.L = tables.L
        self.tables = tables
        self.extents = _f64(extents)
        pd = np.asarray([1 if p else 0 for p in periodic], dtype=np.int32)
        self.h = self.L.l.orc_system_create(tables.h, _ptr(self.extents, _dp), _ptr(pd, _ip))
        self.n = 0

    def rebuild(self, r_cut: float, pos, threads: int = 1) -> None:
        pos = _f64(pos).reshape(-1, 3)
        self.n = len(pos)
        self.L.check(self.L.l.orc_list_rebuild(self.h, r_cut, _ptr(pos, _dp), self.n, threads))

    def update(self, delta) -> None:
        d = _f64(delta).reshape(-1, 3)
        self.L.check(self.L.l.orc_list_update(self.h, _ptr(d, _dp), len(d)))

    def num_plus_ghosts(self) -> int:
        return int(self.L.l.orc_list_num_plus_ghosts(self.h))

    def export(self):
        """(counts[N], real_idx[total], shift_code[total]) of the stored (unfiltered) lists."""
        counts = np.zeros(self.n, dtype=np.int32)
        cap = self.n * 256
        idx = np.zeros(cap, dtype=np.int32)
        sh = np.zeros(cap, dtype=np.int32)
        tot = self.L.l.orc_list_export(self.h, _ptr(counts, _ip), _ptr(idx, _ip), _ptr(sh, _ip), cap)
        assert tot <= cap
        return counts, idx[:tot].copy(), sh[:tot].copy()

    def cells(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.int32)
        self.L.l.orc_list_cells(self.h, _ptr(out, _ip))
        return out

    def for_neighbours(self, i: int, r_cut: float):
        cap = 1024
        idx = np.zeros(cap, dtype=np.int32)
        r = np.zeros(cap)
        dr = np.zeros((cap, 3))
        n = self.L.l.orc_list_for_neighbours(self.h, i, r_cut, _ptr(idx, _ip), _ptr(r, _dp), _ptr(dr, _dp), cap)
        return idx[:n].copy(), r[:n].copy(), dr[:n].copy()

    def energy(self, ids, threads: int = 1) -> float:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = C.c_double()
        self.L.check(self.L.l.orc_eam_energy(self.h, _ptr(ids, _u32p), len(ids), threads, C.byref(out)))
        return out.value

    def gradient(self, ids, frozen, threads: int = 1) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        fz = np.ascontiguousarray(frozen, dtype=np.uint8)
        out = np.zeros((len(ids), 3))
        self.L.check(self.L.l.orc_eam_gradient(self.h, _ptr(ids, _u32p), _ptr(fz, _u8p), len(ids), threads, _ptr(out, _dp)))
        return out

    def hessian(self, ids, frozen, threads: int = 1, mass_weight: bool = False) -> np.ndarray:
        """Dense (3N,3N); only the lower triangle is meaningful (hessian.hpp:86-93)."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        fz = np.ascontiguousarray(frozen, dtype=np.uint8)
        n3 = 3 * len(ids)
        out = np.zeros((n3, n3), dtype=np.float64)
        self.L.check(self.L.l.orc_eam_hessian(self.h, _ptr(ids, _u32p), _ptr(fz, _u8p), len(ids), threads, _ptr(out, _dp),
                                              1 if mass_weight else 0))
        return out.T.copy()  # C wrote column-major

    def minimise(self, pos, ids, frozen, threads: int = 1, iopt=None, dopt=None):
        pos = _f64(pos).reshape(-1, 3)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        fz = np.ascontiguousarray(frozen, dtype=np.uint8)
        iopt = LBFGS_IOPT if iopt is None else np.ascontiguousarray(iopt, dtype=np.int32)
        dopt = LBFGS_DOPT if dopt is None else _f64(dopt)
        out_r = np.zeros_like(pos)
        out_g = np.zeros_like(pos)
        ret = C.c_int()
        cnt = np.zeros(4, dtype=np.int64)
        self.L.check(self.L.l.orc_lbfgs_minimise(self.h, _ptr(iopt, _ip), _ptr(dopt, _dp), _ptr(pos, _dp), _ptr(ids, _u32p),
                                                 _ptr(fz, _u8p), len(pos), threads, _ptr(out_r, _dp), _ptr(out_g, _dp),
                                                 C.byref(ret), _ptr(cnt, _lp)))
        return out_r, out_g, bool(ret.value), cnt

    def minimise_batch(self, pos, ids, frozen, threads: int, iopt=None, dopt=None):
        pos = _f64(pos)
        R, n = pos.shape[0], pos.shape[1]
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        fz = np.ascontiguousarray(frozen, dtype=np.uint8)
        iopt = LBFGS_IOPT if iopt is None else np.ascontiguousarray(iopt, dtype=np.int32)
        dopt = LBFGS_DOPT if dopt is None else _f64(dopt)
        out_r = np.zeros_like(pos)
        out_g = np.zeros_like(pos)
        ret = np.zeros(R, dtype=np.int32)
        cnt = np.zeros((R, 4), dtype=np.int64)
        self.L.check(self.L.l.orc_lbfgs_minimise_batch(self.h, _ptr(iopt, _ip), _ptr(dopt, _dp), _ptr(pos, _dp), _ptr(ids, _u32p),
                                                       _ptr(fz, _u8p), n, R, threads, _ptr(out_r, _dp), _ptr(out_g, _dp),
                                                       _ptr(ret, _ip), _ptr(cnt, _lp)))
        return out_r, out_g, ret, cnt

    def find_sp(self, pos, axis, ids, frozen, in_min, hist_sp=None, theta_tol: float = np.pi / 6, count_frozen: int = 0,
                threads: int = 1, iopt=None, dopt=None):
        pos = _f64(pos).reshape(-1, 3)
        axis = _f64(axis).reshape(-1, 3)
        in_min = _f64(in_min).reshape(-1, 3)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        fz = np.ascontiguousarray(frozen, dtype=np.uint8)
        iopt = DIMER_IOPT if iopt is None else np.ascontiguousarray(iopt, dtype=np.int32)
        dopt = DIMER_DOPT if dopt is None else _f64(dopt)
        hist = _f64(hist_sp if hist_sp is not None else np.zeros((0, len(pos), 3)))
        out_r = np.zeros_like(pos)
        out_ax = np.zeros_like(pos)
        code = C.c_int()
        cnt = np.zeros(4, dtype=np.int64)
        self.L.check(self.L.l.orc_dimer_find_sp(self.h, _ptr(iopt, _ip), _ptr(dopt, _dp), _ptr(pos, _dp), _ptr(axis, _dp),
                                                _ptr(ids, _u32p), _ptr(fz, _u8p), len(pos), _ptr(in_min, _dp), _ptr(hist, _dp),
                                                hist.shape[0], theta_tol, count_frozen, threads, _ptr(out_r, _dp),
                                                _ptr(out_ax, _dp), C.byref(code), _ptr(cnt, _lp)))
        return out_r, out_ax, int(code.value), cnt

    def find_sp_batch(self, pos, axis, ids, frozen, in_min, hist_sp=None, theta_tol: float = np.pi / 6, count_frozen: int = 0,
                      threads: int = 1, iopt=None, dopt=None):
        pos = _f64(pos)
        axis = _f64(axis)
        R, n = pos.shape[0], pos.shape[1]
        in_min = _f64(in_min).reshape(-1, 3)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        fz = np.ascontiguousarray(frozen, dtype=np.uint8)
        iopt = DIMER_IOPT if iopt is None else np.ascontiguousarray(iopt, dtype=np.int32)
        dopt = DIMER_DOPT if dopt is None else _f64(dopt)
        hist = _f64(hist_sp if hist_sp is not None else np.zeros((0, n, 3)))
        out_r = np.zeros_like(pos)
        out_ax = np.zeros_like(pos)
        code = np.zeros(R, dtype=np.int32)
        cnt = np.zeros((R, 4), dtype=np.int64)
        self.L.check(self.L.l.orc_dimer_find_sp_batch(self.h, _ptr(iopt, _ip), _ptr(dopt, _dp), _ptr(pos, _dp), _ptr(axis, _dp),
                                                      _ptr(ids, _u32p), _ptr(fz, _u8p), n, R, _ptr(in_min, _dp), _ptr(hist, _dp),
                                                      hist.shape[0], theta_tol, count_frozen, threads, _ptr(out_r, _dp),
                                                      _ptr(out_ax, _dp), _ptr(code, _ip), _ptr(cnt, _lp)))
        return out_r, out_ax, code, cnt

    def perturb(self, pos, frozen, centre: int, seed, rep: int, r_pert: float = 4.0, stddev: float = 0.6):
        """Master::perturb (src/saddle/find.cpp:1147-1214) with seed chain seed -> long_jump()^rep."""
        pos = _f64(pos).reshape(-1, 3)
        fz = np.ascontiguousarray(frozen, dtype=np.uint8)
        s = np.asarray(seed, dtype=np.uint64)
        out_r = np.zeros_like(pos)
        out_ax = np.zeros_like(pos)
        self.L.check(self.L.l.orc_perturb(self.h, _ptr(pos, _dp), _ptr(fz, _u8p), len(pos), centre, _ptr(s, _u64p), rep, r_pert,
                                          stddev, _ptr(out_r, _dp), _ptr(out_ax, _dp)))
        return out_r, out_ax

    def time_gradient(self, ids, frozen, delta, threads: int, iters: int) -> float:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        fz = np.ascontiguousarray(frozen, dtype=np.uint8)
        d = _f64(delta)
        out = np.zeros((len(ids), 3))
        return float(self.L.l.orc_time_gradient(self.h, _ptr(ids, _u32p), _ptr(fz, _u8p), len(ids), _ptr(d, _dp), threads, iters,
                                                _ptr(out, _dp)))

    def time_rebuild(self, r_cut: float, pos, threads: int, iters: int) -> float:
        pos = _f64(pos).reshape(-1, 3)
        self.n = len(pos)
        return float(self.L.l.orc_time_rebuild(self.h, r_cut, _ptr(pos, _dp), len(pos), threads, iters))
