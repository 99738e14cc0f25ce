"""Host-side mirror of the reference's class surface for the force-evaluation path, on top of the C ABI.

Names, argument meaning and error behaviour follow the reference (paths relative to the openFLY tree):

    Box              system::Box                include/libfly/system/box.hpp:84-197 (orthorhombic only)
    DataEAM          potential::DataEAM         include/libfly/potential/EAM/data.hpp:59-116
    Generic          potential::Generic (EAM)   include/libfly/potential/generic.hpp:79-136
    List             neigh::List                include/libfly/neigh/list.hpp:86-161
    LBFGS            minimise::LBFGS            include/libfly/minimise/LBFGS/lbfgs.hpp:96-118
    Dimer            saddle::Dimer              include/libfly/saddle/dimer.hpp:107-155

The one deliberate extension: every object works on a *batch* of R replicas of the same system
([R, N, 3] arrays); R = 1 with [N, 3] arrays is the reference's single-system call. The C++20 shim with
the same names for a compiled host is include/openfly_b200/fly.hpp.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import DimerOptions, LbfgsOptions, MechOptions, MechRecord, OflyError, OutOfBounds, RotorOptions, check, f64, ptr  # noqa: F401

_dp, _ip, _lp, _u32p, _u8p = capi._dp, capi._ip, capi._lp, capi._u32p, capi._u8p


class Context:
    """One device + stream (there is no global state; one Context per host thread and device)."""

    def __init__(self, device: int = 0):
        self.lib = capi.load()
        h = C.c_void_p()
        check(self.lib.ofly_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def stream(self) -> int:
        return int(self.lib.ofly_ctx_stream(self.h) or 0)

    def synchronize(self) -> None:
        check(self.lib.ofly_ctx_synchronize(self.h))

    def launch_count(self) -> int:
        return int(self.lib.ofly_ctx_launch_count(self.h))

    def tick_count(self) -> int:
        return int(self.lib.ofly_ctx_tick_count(self.h))

    def close(self) -> None:
        if self.h:
            self.lib.ofly_ctx_destroy(self.h)
            self.h = None

    # raw device buffers (bench.py keeps its inputs resident with these)
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(self.lib.ofly_dev_alloc(self.h, nbytes, C.byref(p)))
        return int(p.value)

    def free(self, dev: int) -> None:
        check(self.lib.ofly_dev_free(self.h, C.c_void_p(dev)))

    def upload(self, dev: int, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a)
        check(self.lib.ofly_dev_upload(self.h, C.c_void_p(dev), a.ctypes.data_as(C.c_void_p), a.nbytes))

    def download(self, a: np.ndarray, dev: int) -> None:
        assert a.flags["C_CONTIGUOUS"]
        check(self.lib.ofly_dev_download(self.h, a.ctypes.data_as(C.c_void_p), C.c_void_p(dev), a.nbytes))


@dataclass
class Box:
    """Orthorhombic system::Box: extents (diagonal of the basis) and per-axis periodicity."""

    extents: tuple
    periodic: tuple = (True, True, True)

    @staticmethod
    def from_basis(basis, periodic=(True, True, True)) -> "Box":
        b = np.asarray(basis, dtype=np.float64)
        skew = b - np.diag(np.diag(b))
        if np.sqrt((skew * skew).sum()) >= 1e-4:  # box.hpp:63-76 would build a Triclinic cell
            raise OflyError(capi.OFLY_ERR_INVALID, "triclinic boxes are outside the B200 path (orthorhombic only)")
        return Box(tuple(np.diag(b)), tuple(bool(p) for p in periodic))


class DataEAM:
    """Parsed + splined setfl tables resident on the device (shared, immutable, ref-counted)."""

    def __init__(self, ctx: Context, path: str | None = None, *, text: bytes | None = None, _handle=None):
        """DataEAM(path) reads a setfl file; DataEAM(text=...) parses the same from memory (the reference's
        std::istream constructor, data.cpp:67); from_knots() takes already-splined tables."""
        self.ctx = ctx
        h = C.c_void_p()
        if _handle is not None:
            h = _handle
        elif text is not None:
            check(ctx.lib.ofly_tables_load_setfl_mem(ctx.h, text, len(text), C.byref(h)))
        else:
            check(ctx.lib.ofly_tables_load_setfl(ctx.h, path.encode(), C.byref(h)))
        self.h = h
        info = np.zeros(3, dtype=np.int64)
        dinfo = np.zeros(3 + 8, dtype=np.float64)
        check(ctx.lib.ofly_tables_info(self.h, ptr(info, _lp), ptr(dinfo, _dp)))
        self.n_types, self.knots_rho, self.knots_r = (int(x) for x in info)
        self.d_rho, self.d_r, self._rcut = (float(x) for x in dinfo[:3])
        self.mass = dinfo[3:3 + self.n_types].copy()
        self.names = [ctx.lib.ofly_tables_type_name(self.h, i).decode() for i in range(self.n_types)]

    @classmethod
    def from_knots(cls, ctx: Context, names, mass, d_rho: float, d_r: float, r_cut: float, F, phi, V) -> "DataEAM":
        """F: [n, kF, 4], phi: [n*n, kR, 4] (slot src + n*dst), V: [n*n, kR, 4] (slot max + n*min) spline knots {a,b,c,d}."""
        F, phi, V = f64(F), f64(phi), f64(V)
        n = len(names)
        arr = (C.c_char_p * n)(*[x.encode() for x in names])
        m = f64(mass)
        h = C.c_void_p()
        check(ctx.lib.ofly_tables_create_from_knots(ctx.h, n, arr, ptr(m, _dp), F.shape[1], d_rho, phi.shape[1], d_r, r_cut,
                                                    ptr(F, _dp), ptr(phi, _dp), ptr(V, _dp), C.byref(h)))
        return cls(ctx, _handle=h)

    def check_type_map(self, names) -> None:
        """EAM's constructor check (eam.hpp:50-63): raises unless `names` starts with the table's type names."""
        arr = (C.c_char_p * len(names))(*[x.encode() for x in names])
        check(self.ctx.lib.ofly_eam_check_type_map(self.h, len(names), arr))

    def rcut(self) -> float:
        return self._rcut

    def knots(self, kind: int, i: int, j: int = 0) -> np.ndarray:
        n = self.knots_rho if kind == 0 else self.knots_r
        out = np.zeros((n, 4), dtype=np.float64)
        check(self.ctx.lib.ofly_tables_knots(self.h, kind, i, j, ptr(out, _dp)))
        return out

    def close(self) -> None:
        if self.h:
            self.ctx.lib.ofly_tables_release(self.h)
            self.h = None


class System:
    """R replicas of one N-atom system: the device-side state shared by List / Generic / LBFGS / Dimer."""

    def __init__(self, data: DataEAM, box: Box, type_id, frozen=None, replicas: int = 1):
        self.ctx, self.data, self.box = data.ctx, data, box
        ids = np.ascontiguousarray(type_id, dtype=np.uint32)
        fz = np.zeros(len(ids), dtype=np.uint8) if frozen is None else np.ascontiguousarray(frozen, dtype=np.uint8)
        ex = f64(box.extents)
        pd = np.asarray([1 if p else 0 for p in box.periodic], dtype=np.int32)
        h = C.c_void_p()
        check(self.ctx.lib.ofly_batch_create(self.ctx.h, data.h, ptr(ex, _dp), ptr(pd, _ip), len(ids), replicas, ptr(ids, _u32p),
                                             ptr(fz, _u8p), C.byref(h)))
        self.h = h
        self.N, self.R = len(ids), replicas
        self.type_id, self.frozen = ids, fz

    def _vec(self, a) -> np.ndarray:
        a = f64(a)
        if a.size != self.R * self.N * 3:
            raise OflyError(capi.OFLY_ERR_INVALID, f"expected {self.R}x{self.N}x3 values, got {a.size}")
        return a.reshape(self.R, self.N, 3)

    def _queue(self, a) -> np.ndarray:
        """[Q, N, 3] start states of a queue of searches (any Q >= 1) that run through this system's R slots."""
        a = f64(a)
        if a.size == 0 or a.size % (self.N * 3):
            raise OflyError(capi.OFLY_ERR_INVALID, f"expected Qx{self.N}x3 values, got {a.size}")
        return a.reshape(-1, self.N, 3)

    def close(self) -> None:
        if self.h:
            self.ctx.lib.ofly_batch_destroy(self.h)
            self.h = None


class List:
    """neigh::List(box, r_cut): rebuild / update / for_neighbours over a System's replicas."""

    def __init__(self, system: System, r_cut: float):
        self.sys, self.r_cut = system, float(r_cut)

    def size(self) -> int:
        return self.sys.N

    def rebuild(self, positions, num_threads: int = 1) -> None:
        if num_threads <= 0:
            raise OflyError(capi.OFLY_ERR_INVALID, f"{num_threads} is not a valid number of threads for rebuild()")
        p = self.sys._vec(positions)
        check(self.sys.ctx.lib.ofly_neigh_rebuild(self.sys.h, self.r_cut, ptr(p, _dp)))
        self.sys._list_owner = self  # the System holds ONE device list: whoever rebuilt last owns it (see Generic._owned)

    def update(self, delta) -> None:
        Generic._owned(self)
        d = self.sys._vec(delta)
        check(self.sys.ctx.lib.ofly_neigh_update(self.sys.h, ptr(d, _dp)))

    def export(self, replica: int = 0):
        """(counts[N], real index[total], shift code[total]) of the stored lists."""
        s = self.sys
        counts = np.zeros(s.N, dtype=np.int32)
        cap = s.N * 512
        idx = np.zeros(cap, dtype=np.int32)
        sh = np.zeros(cap, dtype=np.int32)
        tot = C.c_long()
        check(s.ctx.lib.ofly_neigh_export(s.h, replica, ptr(counts, _ip), ptr(idx, _ip), ptr(sh, _ip), cap, C.byref(tot)))
        assert tot.value <= cap
        return counts, idx[:tot.value].copy(), sh[:tot.value].copy()

    def cells(self, replica: int = 0) -> np.ndarray:
        out = np.zeros(self.sys.N, dtype=np.int32)
        check(self.sys.ctx.lib.ofly_neigh_cells(self.sys.h, replica, ptr(out, _ip)))
        return out

    def positions(self, replica: int = 0) -> np.ndarray:
        out = np.zeros((self.sys.N, 3), dtype=np.float64)
        check(self.sys.ctx.lib.ofly_neigh_positions(self.sys.h, replica, ptr(out, _dp)))
        return out

    def mean_count(self) -> float:
        m = C.c_double()
        check(self.sys.ctx.lib.ofly_neigh_mean_count(self.sys.h, C.byref(m)))
        return m.value

    def for_neighbours(self, i: int, r_cut: float, f, replica: int = 0) -> None:
        """Call f(n, r, dr) for every neighbour of atom i within r_cut (list.hpp:138-154). Evaluated on the device for this
        one atom (ofly_neigh_for_atom): only the hits travel."""
        if r_cut > self.r_cut:
            raise OflyError(capi.OFLY_ERR_INVALID, f"r_cut={r_cut} is bigger then than the internal r_cut={self.r_cut}")
        idx, r, dr = self.neighbours_of(i, r_cut, replica)
        for k in range(len(idx)):
            f(int(idx[k]), float(r[k]), dr[k])

    def neighbours_within(self, r_cut: float, replica: int = 0):
        """for_neighbours of every atom in one device pass: (offsets[N+1], index[total], r[total], dr[total, 3]) in CSR form."""
        lib, s = self.sys.ctx.lib, self.sys
        if r_cut > self.r_cut:
            raise OflyError(capi.OFLY_ERR_INVALID, f"r_cut={r_cut} is bigger then than the internal r_cut={self.r_cut}")
        off = np.zeros(s.N + 1, dtype=np.int64)
        tot = C.c_long()
        check(lib.ofly_neigh_within_all(s.h, replica, r_cut, ptr(off, _lp), None, None, None, 0, C.byref(tot)))
        n = max(tot.value, 1)
        idx, r, dr = np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros((n, 3))
        check(lib.ofly_neigh_within_all(s.h, replica, r_cut, ptr(off, _lp), ptr(idx, _ip), ptr(r, _dp), ptr(dr, _dp), n, C.byref(tot)))
        return off, idx[:tot.value], r[:tot.value], dr[:tot.value]

    def neighbours_of(self, i: int, r_cut: float, replica: int = 0):
        """(real index[n], r[n], dr[n, 3]) of atom i's neighbours within r_cut, in stored order."""
        cap = 1024
        idx = np.zeros(cap, dtype=np.int32)
        r = np.zeros(cap)
        dr = np.zeros((cap, 3))
        n = C.c_long()
        check(self.sys.ctx.lib.ofly_neigh_for_atom(self.sys.h, replica, i, r_cut, ptr(idx, _ip), ptr(r, _dp), ptr(dr, _dp), cap, C.byref(n)))
        assert n.value <= cap
        return idx[:n.value].copy(), r[:n.value].copy(), dr[:n.value].copy()


class Generic:
    """potential::Generic holding an EAM: r_cut / energy / gradient / hessian over a List."""

    def __init__(self, data: DataEAM, type_map=None):
        """type_map: the caller's type names by id (system::TypeMap); checked against the table as EAM's constructor
        does (eam.hpp:50-63) when given."""
        self.data = data
        if type_map is not None:
            data.check_type_map(list(type_map))

    def r_cut(self) -> float:
        return self.data.rcut()

    @staticmethod
    def _owned(nl: List) -> None:
        """A System holds one device list shared by every List / LBFGS / Dimer made on it (unlike the reference, where each object
        owns its own; the C++ surface include/openfly_b200/libfly.hpp does own one per object). Evaluating a potential over a
        List whose device state another List or a search has replaced since would silently use the wrong positions: refuse."""
        if getattr(nl.sys, "_list_owner", None) is not nl:
            raise OflyError(capi.OFLY_ERR_INVALID, "this List's device state was replaced (another List.rebuild or a minimise / find_sp "
                                                   "call on the same System): rebuild it first")

    def energy(self, nl: List, num_threads: int = 1) -> np.ndarray:
        self._owned(nl)
        out = np.zeros(nl.sys.R, dtype=np.float64)
        check(nl.sys.ctx.lib.ofly_eam_energy(nl.sys.h, ptr(out, _dp)))
        return out

    def gradient(self, nl: List, num_threads: int = 1) -> np.ndarray:
        self._owned(nl)
        out = np.zeros((nl.sys.R, nl.sys.N, 3), dtype=np.float64)
        check(nl.sys.ctx.lib.ofly_eam_gradient(nl.sys.h, ptr(out, _dp)))
        return out

    def hessian(self, nl: List, replica: int = 0, mass_weight: bool = False, num_threads: int = 1) -> np.ndarray:
        """Dense (3N, 3N) array; only the lower triangle is meaningful (system::Hessian)."""
        self._owned(nl)
        n3 = 3 * nl.sys.N
        out = np.zeros((n3, n3), dtype=np.float64)
        check(nl.sys.ctx.lib.ofly_eam_hessian(nl.sys.h, replica, 1 if mass_weight else 0, ptr(out, _dp)))
        return out.T.copy()  # the library writes column-major

    def hessian_batch(self, nl: List, mass_weight: bool = False) -> np.ndarray:
        """[R, 3N, 3N]: the Hessians of every replica from one launch (lower triangles meaningful)."""
        self._owned(nl)
        n3 = 3 * nl.sys.N
        out = np.zeros((nl.sys.R, n3, n3), dtype=np.float64)
        check(nl.sys.ctx.lib.ofly_eam_hessian_batch(nl.sys.h, 1 if mass_weight else 0, ptr(out, _dp)))
        return np.ascontiguousarray(out.transpose(0, 2, 1))  # the library writes column-major

    def hessian_eigenvalues(self, nl: List, replica: int = 0, mass_weight: bool = True) -> np.ndarray:
        """system::Hessian::eigenvalues() of the (mass-weighted) Hessian, solved on the device: only 3N doubles return."""
        self._owned(nl)
        out = np.zeros(3 * nl.sys.N, dtype=np.float64)
        check(nl.sys.ctx.lib.ofly_eam_hessian_eigenvalues(nl.sys.h, replica, 1 if mass_weight else 0, ptr(out, _dp)))
        return out


def lbfgs_options(**kw) -> LbfgsOptions:
    o = LbfgsOptions()
    capi.load().ofly_lbfgs_options_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def dimer_options(**kw) -> DimerOptions:
    o = DimerOptions()
    capi.load().ofly_dimer_options_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def rotor_options(**kw) -> RotorOptions:
    o = RotorOptions()
    capi.load().ofly_rotor_options_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class LBFGS:
    """minimise::LBFGS(opt, box). minimise() returns (positions, gradient, failed[R], counters[R,4]);
    failed follows the reference's bool: False on convergence, True on iter_max. A replica that hit
    Spline::OutOfBounds raises OutOfBounds (the reference throws) unless raise_oob=False, in which case
    its status is left at -5 in `self.status`."""

    def __init__(self, opt: LbfgsOptions | None = None):
        self.opt = opt or lbfgs_options()
        self.status = None

    def minimise(self, system: System, positions, pot: Generic, num_threads: int = 1, raise_oob: bool = True):
        """positions: [Q, N, 3]; Q == system.R is the reference's batch, any other Q streams through the R slots."""
        p = system._queue(positions)
        Q = p.shape[0]
        system._list_owner = None  # the search rebuilds the System's device list at its own radius
        out_r = np.zeros_like(p)
        out_g = np.zeros_like(p)
        status = np.zeros(Q, dtype=np.int32)
        cnt = np.zeros((Q, 4), dtype=np.int64)
        check(system.ctx.lib.ofly_lbfgs_minimise_stream(system.h, C.byref(self.opt), Q, ptr(p, _dp), ptr(out_r, _dp), ptr(out_g, _dp),
                                                        ptr(status, _ip), ptr(cnt, _lp)))
        self.status = status
        if raise_oob and (status == -5).any():
            raise OutOfBounds(capi.OFLY_ERR_OUT_OF_BOUNDS, "x is outside tabulated region")
        return out_r, out_g, status != 0, cnt


class Dimer:
    """saddle::Dimer(opt, rotor opt, box). find_sp() returns (positions, axis, exit[Q], counters[Q,4]) with
    exit codes Dimer::Exit (0 success, -1 convex, -2 iter_max, -3 collision); -5 marks OutOfBounds. positions / axis
    are [Q, N, 3]: Q == system.R is the reference's batch; any other Q is a queue that streams through the R slots
    (a finished slot takes the next start state on the device), with the same per-search results."""

    success, convex, iter_max, collision, uninit = 0, -1, -2, -3, -4

    def __init__(self, opt: DimerOptions | None = None, r_opt: RotorOptions | None = None):
        self.opt = opt or dimer_options()
        self.r_opt = r_opt or rotor_options()

    def find_sp(self, system: System, positions, axis, in_min, pot: Generic, hist_sp=None, theta_tol: float = np.pi / 6,
                count_frozen: int = 0, num_threads: int = 1, raise_oob: bool = False):
        p = system._queue(positions)
        ax = system._queue(axis)
        Q = p.shape[0]
        system._list_owner = None  # the search rebuilds the System's device list at its own radius
        if ax.shape != p.shape:
            raise OflyError(capi.OFLY_ERR_INVALID, "positions and axes differ in shape")
        mn = f64(in_min).reshape(system.N, 3)
        hist = f64(np.zeros((0, system.N, 3)) if hist_sp is None else hist_sp).reshape(-1, system.N, 3)
        out_r = np.zeros_like(p)
        out_ax = np.zeros_like(p)
        code = np.zeros(Q, dtype=np.int32)
        cnt = np.zeros((Q, 4), dtype=np.int64)
        check(system.ctx.lib.ofly_dimer_find_sp_stream(system.h, C.byref(self.opt), C.byref(self.r_opt), Q, ptr(p, _dp), ptr(ax, _dp),
                                                       ptr(mn, _dp), ptr(hist, _dp), hist.shape[0], theta_tol, count_frozen,
                                                       ptr(out_r, _dp), ptr(out_ax, _dp), ptr(code, _ip), ptr(cnt, _lp)))
        if raise_oob and (code == -5).any():
            raise OutOfBounds(capi.OFLY_ERR_OUT_OF_BOUNDS, "x is outside tabulated region")
        return out_r, out_ax, code, cnt


def mech_options(**kw) -> MechOptions:
    o = MechOptions()
    capi.load().ofly_mech_options_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Mechanisms:
    """The numerics of saddle::Master that follow the dimer searches of one environment (src/saddle/find.cpp), batched on
    the device: calc_minima_hess (:1094-1144), m_sep_list (:223-244), do_adj_min (:1014-1090), the energy / Hessian /
    eigenvalue half of saddle_2_mech (:757-871) and the prefactor (:287). Thin callers of the ofly_mech_* C entry points."""

    FIELDS = ("status", "count_zeros", "i2r", "r2f", "r2w", "w2f", "E0", "Esp", "Ef", "barrier", "delta", "log_prod_eigen", "min_mode")

    def __init__(self, data: DataEAM, box: Box, type_id, frozen=None, replicas: int = 64):
        self.sys = System(data, box, type_id, frozen, replicas=replicas)
        self.lib = data.ctx.lib
        self.log_prod_eigen = None
        self.num_zero_modes = None

    def close(self) -> None:
        self.sys.close()

    def perturb(self, in_min, centre: int, stddev, normals, r_pert: float = 4.0):
        """Master::perturb (find.cpp:1147-1214) for S start states: stddev [S], normals [S, N, 6] (the caller's random draws).
        Returns (positions [S, N, 3], unit axes [S, N, 3])."""
        N = self.sys.N
        mn = f64(in_min).reshape(N, 3)
        sd = f64(stddev).reshape(-1)
        z = f64(normals).reshape(len(sd), N, 6)
        pos = np.zeros((len(sd), N, 3))
        ax = np.zeros_like(pos)
        check(self.lib.ofly_master_perturb(self.sys.h, len(sd), ptr(mn, _dp), centre, r_pert, ptr(sd, _dp), ptr(z, _dp), ptr(pos, _dp), ptr(ax, _dp)))
        return pos, ax

    def skmc_relax_step(self, cell, raw_recon, min_opt: LbfgsOptions | None = None):
        """The numeric part of one SKMC step (skmc.hpp:297-309): returns dict(E0, Ef, dR, failed, rel_recon)."""
        N = self.sys.N
        a, b = f64(cell).reshape(N, 3), f64(raw_recon).reshape(N, 3)
        rel = np.zeros((N, 3))
        e0, ef, dr, failed = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        o = min_opt or lbfgs_options()
        check(self.lib.ofly_skmc_relax_step(self.sys.h, C.byref(o), ptr(a, _dp), ptr(b, _dp), ptr(rel, _dp), C.byref(e0), C.byref(ef), C.byref(dr),
                                            C.byref(failed)))
        return {"E0": e0.value, "Ef": ef.value, "dR": dr.value, "failed": bool(failed.value), "rel_recon": rel}

    def sep_list(self, positions) -> np.ndarray:
        p = f64(positions).reshape(self.sys.N, 3)
        out = np.zeros(self.sys.N, dtype=np.int32)
        check(self.lib.ofly_mech_sep_list(self.sys.h, ptr(p, _dp), ptr(out, _ip)))
        return out

    def minimum(self, in_min, zero_tol: float = 1e-6):
        """calc_minima_hess: returns (sum log lambda, zero modes, frozen atoms, eigenvalues[3N]) and remembers the first two."""
        p = f64(in_min).reshape(self.sys.N, 3)
        lp, nz, nf = C.c_double(), C.c_int(), C.c_int()
        eig = np.zeros(3 * self.sys.N)
        check(self.lib.ofly_mech_minimum(self.sys.h, ptr(p, _dp), zero_tol, C.byref(lp), C.byref(nz), C.byref(nf), ptr(eig, _dp)))
        self.log_prod_eigen, self.num_zero_modes = lp.value, nz.value
        return lp.value, nz.value, nf.value, eig

    def _run(self, opt, lopt, in_min, sp, axis, want_path):
        N = self.sys.N
        p = f64(sp).reshape(-1, N, 3)
        S = p.shape[0]
        ax = None if axis is None else f64(axis).reshape(S, N, 3)
        mn = f64(in_min).reshape(N, 3)
        rec = (MechRecord * S)()
        outs = [np.zeros_like(p) for _ in range(3)] if want_path else [None] * 3
        arg = [ptr(o, _dp) if o is not None else None for o in outs]
        check(self.lib.ofly_mech_from_saddles(self.sys.h, C.byref(opt), C.byref(lopt), S, ptr(p, _dp), None if ax is None else ptr(ax, _dp),
                                              ptr(mn, _dp), arg[0], arg[1], arg[2], rec))
        res = {k: np.array([getattr(r, k) for r in rec]) for k in self.FIELDS}
        if want_path:
            res["rev"], res["sp"], res["fwd"] = outs
        return res

    def from_saddles(self, in_min, sp, axis, centre: int = -1, freeze_atom: int = -1, opt: MechOptions | None = None,
                     min_opt: LbfgsOptions | None = None):
        """do_adj_min + saddle_2_mech numerics for S dimer results (sp, axis: [S, N, 3]). Returns a dict of per-saddle arrays
        (ofly_mech_record fields, `prefactor` in Hz for status 0) plus the drift-corrected pathway rev / sp / fwd."""
        if self.num_zero_modes is None:
            self.minimum(in_min)
        o = opt or mech_options()
        o.centre, o.freeze_atom, o.num_zero_modes, o.skip_pathway = centre, freeze_atom, self.num_zero_modes, 0
        res = self._run(o, min_opt or lbfgs_options(), in_min, sp, axis, True)
        res["prefactor"] = np.array([self.lib.ofly_mech_prefactor(self.log_prod_eigen, lp) if st == 0 else np.nan
                                     for st, lp in zip(res["status"], res["log_prod_eigen"])])
        return res

    def prefactors(self, in_min, sp, opt: MechOptions | None = None):
        """Only the Hessian half: barrier against in_min, eigenvalue rules and prefactor of every saddle in sp [S, N, 3].
        `valid` is 1 where the saddle is first order with the minimum's zero-mode count."""
        if self.num_zero_modes is None:
            self.minimum(in_min)
        o = opt or mech_options()
        o.num_zero_modes, o.skip_pathway = self.num_zero_modes, 1
        res = self._run(o, lbfgs_options(), in_min, sp, None, False)
        res["valid"] = (res["status"] == 0).astype(np.int32)
        res["prefactor"] = np.array([self.lib.ofly_mech_prefactor(self.log_prod_eigen, lp) if st == 0 else np.nan
                                     for st, lp in zip(res["status"], res["log_prod_eigen"])])
        return res


# ---- Vineyard prefactor from eigenvalues on the host (numpy): the checker of ofly_mech_* in the tests ----

def hessian_eigenvalues(hess_lower: np.ndarray) -> np.ndarray:
    """system::Hessian::eigenvalues (hessian.hpp:86-93): ascending eigenvalues of the symmetric matrix whose LOWER
    triangle is given (Eigen's SelfAdjointEigenSolver reads the lower triangle only)."""
    h = np.tril(hess_lower)
    return np.linalg.eigvalsh(h + np.tril(h, -1).T)


def log_prod_eigen_minimum(eigs: np.ndarray, count_frozen: int = 0, zero_tol: float = 1e-6):
    """Master::calc_minima_hess (find.cpp:1105-1135): sum of log(lambda) over the non-zero modes of the minimum; every mode
    must be > -zero_tol and the number of zero modes must be 3 (3 * count_frozen when atoms are frozen)."""
    if (eigs < -zero_tol).any():
        raise OflyError(capi.OFLY_ERR_INVALID, f"Mode has eigenvalue = {eigs.min()}")
    nz = eigs > zero_tol
    zeros = int((~nz).sum())
    expect = 3 * count_frozen if count_frozen > 0 else 3
    if zeros != expect:
        raise OflyError(capi.OFLY_ERR_INVALID, f"Expecting {expect} zero modes but fount {zeros}")
    return float(np.log(eigs[nz]).sum()), zeros


def vineyard_prefactor(eigs_min: np.ndarray, eigs_sp: np.ndarray, count_frozen: int = 0, zero_tol: float = 1e-6):
    """Harmonic-TST prefactor as saddle::Master computes it (find.cpp:284-290, 846-871): the saddle needs exactly one
    negative mode (index 0), the same number of zero modes as the minimum, and
        nu = sqrt(exp(sum log lambda_min - sum log lambda_sp) / (2 pi 1.6605390666050e-27)).
    Returns None for a second-or-higher-order saddle (the reference discards the mechanism)."""
    log_min, zeros_min = log_prod_eigen_minimum(eigs_min, count_frozen, zero_tol)
    if not eigs_sp[0] < -zero_tol:
        raise OflyError(capi.OFLY_ERR_INVALID, f"Saddle-point with minimum mode={eigs_sp[0]}")
    rest = eigs_sp[1:]
    if (rest < -zero_tol).any():
        return None
    nz = rest > zero_tol
    zeros = int((~nz).sum())
    if zeros != zeros_min:
        raise OflyError(capi.OFLY_ERR_INVALID, f"Expecting {zeros_min} zero modes but hessian has {zeros}")
    log_sp = float(np.log(rest[nz]).sum())
    return float(np.sqrt(np.exp(log_min - log_sp) / (2 * np.pi * 1.6605390666050e-27)))
