"""Host-side restatement of the prefactor bookkeeping of saddle::Master (find.cpp:284-290, 846-871, 1105-1135): no GPU."""
import numpy as np
import pytest


def test_vineyard_prefactor_bookkeeping():
    from openfly_b200 import fly

    rng = np.random.default_rng(0)
    pos_modes = rng.uniform(0.02, 20.0, 30)
    ev_min = np.sort(np.concatenate([[1e-15, -2e-16, 3e-16], pos_modes]))
    sp_modes = pos_modes[1:] * rng.uniform(0.9, 1.1, 29)
    ev_sp = np.sort(np.concatenate([[-0.5], [1e-16, 0.0, -1e-15], sp_modes]))
    nu = fly.vineyard_prefactor(ev_min, ev_sp)
    expect = np.sqrt(np.exp(np.log(pos_modes).sum() - np.log(sp_modes).sum()) / (2 * np.pi * 1.6605390666050e-27))
    assert abs(nu / expect - 1) < 1e-12
    # the negative mode itself never enters the product (the loop starts at index 1, find.cpp:855)
    ev_sp2 = ev_sp.copy()
    ev_sp2[0] = -5.0
    assert fly.vineyard_prefactor(ev_min, ev_sp2) == nu
    # a second negative mode: the reference drops the mechanism
    ev_sp3 = np.sort(np.concatenate([[-0.5, -0.01], [0.0, 0.0, 0.0], sp_modes[1:]]))
    assert fly.vineyard_prefactor(ev_min, ev_sp3) is None
    # no negative mode at the saddle / a negative mode at the minimum / wrong number of zero modes: errors
    with pytest.raises(fly.OflyError):
        fly.vineyard_prefactor(ev_min, np.sort(np.abs(ev_sp)))
    with pytest.raises(fly.OflyError):
        fly.vineyard_prefactor(np.concatenate([[-0.1], ev_min[1:]]), ev_sp)
    with pytest.raises(fly.OflyError):
        fly.vineyard_prefactor(ev_min[1:], ev_sp)
    # frozen atoms: 3 zero modes per frozen atom (find.cpp:1101)
    ev_fz = np.sort(np.concatenate([np.zeros(6), pos_modes]))
    assert fly.log_prod_eigen_minimum(ev_fz, count_frozen=2)[1] == 6


def test_hessian_eigenvalues_reads_lower_triangle_only():
    from openfly_b200 import fly

    rng = np.random.default_rng(1)
    a = rng.normal(size=(12, 12))
    sym = a + a.T
    lower = np.tril(sym) + np.triu(rng.normal(size=(12, 12)), 1)  # garbage above the diagonal, as system::Hessian leaves it
    assert np.allclose(fly.hessian_eigenvalues(lower), np.linalg.eigvalsh(sym), atol=1e-12)
