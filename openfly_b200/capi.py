"""ctypes binding of the C ABI declared in include/openfly_b200.h.

This is the only way Python code in this repository reaches the CUDA path: tests, bench.py and
__graft_entry__ all call the same `extern "C"` symbols a C/C++ host program would link against.
There is no CPU fallback: if the shared library is missing or the machine has no GPU the calls fail.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OFLY_LIB") or os.path.join(_PKG, "libopenfly_b200.so")  # OFLY_LIB: a tuning build (csrc/Makefile VARIANT=)
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "openfly_b200.h")

OFLY_OK, OFLY_ERR_INVALID, OFLY_ERR_OUT_OF_BOUNDS, OFLY_ERR_CUDA, OFLY_ERR_CAPACITY = range(5)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p


class OflyError(RuntimeError):
    """fly::RuntimeError (include/libfly/utility/core.hpp:89-138)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[ofly status {code}] {msg}")
        self.code = code


class OutOfBounds(OflyError):
    """fly::Spline::OutOfBounds (include/libfly/utility/spline.hpp:46-51)."""


class LbfgsOptions(C.Structure):
    _fields_ = [("n", C.c_int), ("iter_max", C.c_int), ("f2norm", C.c_double), ("skin_frac", C.c_double), ("proj_tol", C.c_double),
                ("max_trust", C.c_double), ("min_trust", C.c_double), ("grow_trust", C.c_double), ("shrink_trust", C.c_double)]


class DimerOptions(C.Structure):
    _fields_ = [("n", C.c_int), ("max_steps", C.c_int), ("hist_check_freq", C.c_int), ("use_history", C.c_int),
                ("f2norm", C.c_double), ("skin_frac", C.c_double), ("proj_tol", C.c_double), ("max_trust", C.c_double),
                ("min_trust", C.c_double), ("grow_trust", C.c_double), ("shrink_trust", C.c_double), ("convex_max", C.c_double)]


class RotorOptions(C.Structure):
    _fields_ = [("n", C.c_int), ("iter_max_rot", C.c_int), ("relax_in_convex", C.c_int), ("delta_r", C.c_double),
                ("theta_tol", C.c_double)]


class MechOptions(C.Structure):
    _fields_ = [("nudge_frac", C.c_double), ("basin_tol", C.c_double), ("stationary_tol", C.c_double),
                ("hessian_eigen_zero_tol", C.c_double), ("centre", C.c_long), ("freeze_atom", C.c_long), ("num_zero_modes", C.c_int),
                ("skip_pathway", C.c_int)]


class MechRecord(C.Structure):
    _fields_ = [("status", C.c_int), ("count_zeros", C.c_int), ("i2r", C.c_double), ("r2f", C.c_double), ("r2w", C.c_double),
                ("w2f", C.c_double), ("E0", C.c_double), ("Esp", C.c_double), ("Ef", C.c_double), ("barrier", C.c_double),
                ("delta", C.c_double), ("log_prod_eigen", C.c_double), ("min_mode", C.c_double)]


# symbol -> (restype, argtypes); the test-suite checks this table against the header.
SIGNATURES = {
    "ofly_last_error": (C.c_char_p, []),
    "ofly_version": (C.c_char_p, []),
    "ofly_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ofly_ctx_destroy": (None, [_vp]),
    "ofly_ctx_stream": (_vp, [_vp]),
    "ofly_ctx_synchronize": (C.c_int, [_vp]),
    "ofly_ctx_launch_count": (C.c_long, [_vp]),
    "ofly_ctx_tick_count": (C.c_long, [_vp]),
    "ofly_tables_load_setfl": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "ofly_tables_load_setfl_mem": (C.c_int, [_vp, C.c_char_p, C.c_size_t, C.POINTER(_vp)]),
    "ofly_tables_create_from_knots": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_char_p), _dp, C.c_long, C.c_double, C.c_long, C.c_double,
                                                C.c_double, _dp, _dp, _dp, C.POINTER(_vp)]),
    "ofly_tables_type_name": (C.c_char_p, [_vp, C.c_int]),
    "ofly_eam_check_type_map": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_char_p)]),
    "ofly_tables_retain": (None, [_vp]),
    "ofly_tables_release": (None, [_vp]),
    "ofly_tables_info": (C.c_int, [_vp, _lp, _dp]),
    "ofly_tables_knots": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp]),
    "ofly_setfl_info_host": (C.c_int, [C.c_char_p, _lp, _dp]),
    "ofly_setfl_knots_host": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, _dp]),
    "ofly_eam_r_cut": (C.c_double, [_vp]),
    "ofly_batch_create": (C.c_int, [_vp, _vp, _dp, _ip, C.c_long, C.c_long, _u32p, _u8p, C.POINTER(_vp)]),
    "ofly_batch_destroy": (None, [_vp]),
    "ofly_batch_atoms": (C.c_long, [_vp]),
    "ofly_batch_replicas": (C.c_long, [_vp]),
    "ofly_neigh_rebuild": (C.c_int, [_vp, C.c_double, _dp]),
    "ofly_neigh_rebuild_dev": (C.c_int, [_vp, C.c_double, _vp]),
    "ofly_neigh_update": (C.c_int, [_vp, _dp]),
    "ofly_neigh_update_dev": (C.c_int, [_vp, _vp]),
    "ofly_neigh_export": (C.c_int, [_vp, C.c_long, _ip, _ip, _ip, C.c_long, _lp]),
    "ofly_neigh_cells": (C.c_int, [_vp, C.c_long, _ip]),
    "ofly_neigh_positions": (C.c_int, [_vp, C.c_long, _dp]),
    "ofly_neigh_mean_count": (C.c_int, [_vp, _dp]),
    "ofly_eam_energy": (C.c_int, [_vp, _dp]),
    "ofly_eam_gradient": (C.c_int, [_vp, _dp]),
    "ofly_eam_gradient_dev": (C.c_int, [_vp, _vp]),
    "ofly_eam_hessian": (C.c_int, [_vp, C.c_long, C.c_int, _dp]),
    "ofly_lbfgs_options_default": (None, [C.POINTER(LbfgsOptions)]),
    "ofly_dimer_options_default": (None, [C.POINTER(DimerOptions)]),
    "ofly_rotor_options_default": (None, [C.POINTER(RotorOptions)]),
    "ofly_lbfgs_minimise": (C.c_int, [_vp, C.POINTER(LbfgsOptions), _dp, _dp, _dp, _ip, _lp]),
    "ofly_lbfgs_minimise_dev": (C.c_int, [_vp, C.POINTER(LbfgsOptions), _vp, _vp, _vp, _vp, _vp]),
    "ofly_dimer_find_sp": (C.c_int, [_vp, C.POINTER(DimerOptions), C.POINTER(RotorOptions), _dp, _dp, _dp, _dp, C.c_long,
                                     C.c_double, C.c_int, _dp, _dp, _ip, _lp]),
    "ofly_dimer_find_sp_dev": (C.c_int, [_vp, C.POINTER(DimerOptions), C.POINTER(RotorOptions), _vp, _vp, _vp, _vp, C.c_long,
                                         C.c_double, C.c_int, _vp, _vp, _vp, _vp]),
    "ofly_lbfgs_minimise_stream": (C.c_int, [_vp, C.POINTER(LbfgsOptions), C.c_long, _dp, _dp, _dp, _ip, _lp]),
    "ofly_lbfgs_minimise_stream_dev": (C.c_int, [_vp, C.POINTER(LbfgsOptions), C.c_long, _vp, _vp, _vp, _vp, _vp]),
    "ofly_dimer_find_sp_stream": (C.c_int, [_vp, C.POINTER(DimerOptions), C.POINTER(RotorOptions), C.c_long, _dp, _dp, _dp, _dp,
                                            C.c_long, C.c_double, C.c_int, _dp, _dp, _ip, _lp]),
    "ofly_dimer_find_sp_stream_dev": (C.c_int, [_vp, C.POINTER(DimerOptions), C.POINTER(RotorOptions), C.c_long, _vp, _vp, _vp, _vp,
                                                C.c_long, C.c_double, C.c_int, _vp, _vp, _vp, _vp]),
    "ofly_eam_hessian_eigenvalues": (C.c_int, [_vp, C.c_long, C.c_int, _dp]),
    "ofly_neigh_for_atom": (C.c_int, [_vp, C.c_long, C.c_long, C.c_double, _ip, _dp, _dp, C.c_long, _lp]),
    "ofly_neigh_within_all": (C.c_int, [_vp, C.c_long, C.c_double, _lp, _ip, _dp, _dp, C.c_long, _lp]),
    "ofly_master_perturb": (C.c_int, [_vp, C.c_long, _dp, C.c_long, C.c_double, _dp, _dp, _dp, _dp]),
    "ofly_skmc_relax_step": (C.c_int, [_vp, C.POINTER(LbfgsOptions), _dp, _dp, _dp, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "ofly_eam_hessian_batch": (C.c_int, [_vp, C.c_int, _dp]),
    "ofly_mech_options_default": (None, [C.POINTER(MechOptions)]),
    "ofly_mech_sep_list": (C.c_int, [_vp, _dp, _ip]),
    "ofly_mech_minimum": (C.c_int, [_vp, _dp, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), _dp]),
    "ofly_mech_from_saddles": (C.c_int, [_vp, C.POINTER(MechOptions), C.POINTER(LbfgsOptions), C.c_long, _dp, _dp, _dp, _dp, _dp, _dp,
                                         C.POINTER(MechRecord)]),
    "ofly_mech_prefactor": (C.c_double, [C.c_double, C.c_double]),
    "ofly_bench_fp64_fma": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "ofly_dev_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "ofly_dev_free": (C.c_int, [_vp, _vp]),
    "ofly_dev_upload": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "ofly_dev_download": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
}


def build(verbose: bool = False) -> str:
    """Compile libopenfly_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-j4", "-C", os.path.join(_PKG, "csrc")], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libopenfly_b200.so failed")
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(openfly_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc == OFLY_OK:
        return
    msg = load().ofly_last_error().decode()
    if rc == OFLY_ERR_OUT_OF_BOUNDS:
        raise OutOfBounds(rc, msg)
    raise OflyError(rc, msg)


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)
