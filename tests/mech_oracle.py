"""TEST INFRASTRUCTURE. CPU composition of saddle::Master's per-mechanism numerics from the oracle's primitives, statement by
statement after src/saddle/find.cpp: do_adj_min (:1014-1090), saddle_2_mech up to the eigenvalue rules (:757-871),
calc_minima_hess (:1094-1144), m_sep_list (:223-244), prefactor (:287). Eigenvalues: numpy.linalg.eigvalsh on the oracle's
mass-weighted Hessian (Eigen's SelfAdjointEigenSolver in the reference; any symmetric solver agrees to ~1e-12)."""
import numpy as np

ZERO_TOL = 1e-6


def eigenvalues(hess_lower):
    h = np.tril(hess_lower)
    return np.linalg.eigvalsh(h + np.tril(h, -1).T)


def sep_list(pos, extents):
    out = np.zeros(len(pos), dtype=np.int32)
    for i in range(len(pos)):
        d = pos - pos[i]
        d -= extents * np.rint(d / extents)
        out[i] = int(np.argmax(np.sqrt((d * d).sum(1))))  # first of the maxima, as `dr > max` keeps it
    return out


def minimum(osys, otab, ids, fz, r_min, threads=8):
    osys.rebuild(otab.rcut, r_min, threads)
    freq = eigenvalues(osys.hessian(ids, fz, threads, mass_weight=True))
    assert (freq > -ZERO_TOL).all()
    nz = freq > ZERO_TOL
    return float(np.log(freq[nz]).sum()), int((~nz).sum()), freq


def saddle_to_mech(osys, otab, ids, fz, r_min, sp, axis, centre, freeze_atom, num_zero_modes, nudge_frac=0.01, basin_tol=0.05,
                   stationary_tol=0.025, threads=8):
    """Returns a dict like ofly_mech_record (+ rev / sp / fwd); status codes as include/openfly_b200.h."""
    out = {"status": 0}
    disp = np.sqrt(((sp - r_min) ** 2).sum())
    fz2 = fz.copy()
    if not fz.any():
        fz2[freeze_atom] = 1
    fwd, _, fail_f, _ = osys.minimise(sp + axis * disp * nudge_frac, ids, fz2, threads=threads)
    rev, _, fail_r, _ = osys.minimise(sp - axis * disp * nudge_frac, ids, fz2, threads=threads)
    if fail_f or fail_r:
        out["status"] = 1
        return out
    if np.sqrt(((rev - r_min) ** 2).sum()) > np.sqrt(((fwd - r_min) ** 2).sum()):
        fwd, rev = rev, fwd
    drift = rev.mean(axis=0) - r_min.mean(axis=0)
    fwd, rev, spc = fwd - drift, rev - drift, sp - drift
    out.update(rev=rev, sp=spc, fwd=fwd)
    d = ((rev - fwd) ** 2)
    if int(np.argmax((d[:, 0] + d[:, 1]) + d[:, 2])) != centre:
        out["status"] = 2
        return out
    gn = lambda a: float(np.sqrt((a * a).sum()))  # noqa: E731
    out.update(i2r=gn(rev - r_min), r2f=gn(fwd - rev), r2w=gn(rev - spc), w2f=gn(fwd - spc))
    if out["i2r"] > basin_tol:
        out["status"] = 3
    elif out["r2f"] < basin_tol:
        out["status"] = 4
    elif out["r2w"] < stationary_tol:
        out["status"] = 5
    elif out["w2f"] < stationary_tol:
        out["status"] = 6
    if out["status"]:
        return out

    def energy(r):
        osys.rebuild(otab.rcut, r, threads)
        return osys.energy(ids, threads)

    out["E0"], out["Esp"], out["Ef"] = energy(rev), energy(spc), energy(fwd)
    out["barrier"], out["delta"] = out["Esp"] - out["E0"], out["Ef"] - out["E0"]
    if not out["Esp"] > out["E0"]:
        out["status"] = -1
        return out
    osys.rebuild(otab.rcut, spc, threads)
    freq = eigenvalues(osys.hessian(ids, fz, threads, mass_weight=True))
    out["min_mode"] = float(freq[0])
    if not freq[0] < -ZERO_TOL:
        out["status"] = -2
        return out
    rest = freq[1:]
    if (rest < -ZERO_TOL).any():
        out["status"] = 7
        return out
    nz = rest > ZERO_TOL
    out["count_zeros"] = int((~nz).sum())
    out["log_prod_eigen"] = float(np.log(rest[nz]).sum())
    if out["count_zeros"] != num_zero_modes:
        out["status"] = -3
    return out


def prefactor(log_min, log_sp):
    return float(np.sqrt(np.exp(log_min - log_sp) / (2 * np.pi * 1.6605390666050e-27)))
