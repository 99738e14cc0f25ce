"""C-ABI boundary without a GPU: the library loads, exports every symbol include/openfly_b200.h declares,
its host-side logic (setfl parse + splines, option defaults, error reporting) matches the oracle / the
reference's defaults, and it fails loudly -- never falls back -- when there is no device or no library."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.common import ROOT, WEN

HEADER = os.path.join(ROOT, "include", "openfly_b200.h")


@pytest.fixture(scope="module")
def lib():
    from openfly_b200 import capi

    capi.build()
    return capi.load()


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ofly_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from openfly_b200 import capi

    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/openfly_b200.h but not exported"
    missing = [n for n in names if n not in capi.SIGNATURES and n not in ("ofly_last_error", "ofly_version")]
    assert not missing, f"no ctypes signature for {missing}"


def test_setfl_tables_bit_identical_to_oracle(lib, oracle_lib):
    """DataEAM ctor + Spline ctor (data.cpp:67-153, spline.cpp:27-103) on the host side of the library."""
    tab = oracle_lib.Tables(WEN)
    info = (C.c_long * 3)()
    dinfo = (C.c_double * 11)()
    assert lib.ofly_setfl_info_host(WEN.encode(), info, dinfo) == 0
    assert list(info) == [2, 10004, 10004]  # knots = numP + 4 (five padded copies of the last value)
    assert dinfo[0] == 0.02 and dinfo[1] == 5.3e-4 and dinfo[2] == 5.3
    assert [dinfo[3], dinfo[4]] == [55.847, 1.008]
    for kind, pairs in ((0, [(0, 0), (1, 0)]), (1, [(0, 0), (0, 1), (1, 0), (1, 1)]), (2, [(0, 0), (1, 0), (1, 1)])):
        for i, j in pairs:
            got = np.zeros((10004, 4))
            assert lib.ofly_setfl_knots_host(WEN.encode(), kind, i, j, got.ctypes.data_as(C.POINTER(C.c_double))) == 0
            assert np.array_equal(got, tab.knots(kind, i, j)), (kind, i, j)


def test_option_defaults_are_the_reference_defaults(lib):
    """lbfgs.hpp:53-88, dimer.hpp:56-98, rotor.hpp:47-60"""
    from openfly_b200 import fly

    lo, do, ro = fly.lbfgs_options(), fly.dimer_options(), fly.rotor_options()
    assert (lo.n, lo.iter_max, lo.f2norm, lo.skin_frac, lo.proj_tol, lo.max_trust, lo.min_trust, lo.grow_trust, lo.shrink_trust) == \
        (10, 2000, 1e-6, 1.1, 0.0, 0.5, 0.01, 1.5, 0.5)
    assert (do.n, do.max_steps, do.hist_check_freq, do.use_history, do.f2norm, do.skin_frac, do.proj_tol, do.max_trust, do.min_trust,
            do.grow_trust, do.shrink_trust, do.convex_max) == (10, 1000, 5, 1, 1e-5, 1.1, 0.0, 0.5, 0.05, 1.5, 0.5, 3.0)
    assert (ro.n, ro.iter_max_rot, ro.relax_in_convex, ro.delta_r, ro.theta_tol) == (6, 20, 1, 0.001, np.pi / 360)


def test_errors_are_reported_not_swallowed(lib):
    from openfly_b200 import capi

    info = (C.c_long * 3)()
    dinfo = (C.c_double * 11)()
    rc = lib.ofly_setfl_info_host(b"/nonexistent/file.eam.fs", info, dinfo)
    assert rc == capi.OFLY_ERR_INVALID and b"nonexistent" in lib.ofly_last_error()
    with pytest.raises(capi.OflyError):
        capi.check(rc)


def test_no_device_means_an_error_not_a_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from openfly_b200 import capi, fly

    with pytest.raises(capi.OflyError) as e:
        fly.Context(0)
    assert e.value.code == capi.OFLY_ERR_CUDA


def test_missing_library_raises(monkeypatch):
    from openfly_b200 import capi

    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", os.path.join(ROOT, "openfly_b200", "does_not_exist.so"))
    with pytest.raises(FileNotFoundError):
        capi.load()
