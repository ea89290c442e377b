"""ctypes binding of libgeordd_b200.so (include/geordd_b200.h).

This is the Python twin of the Julia `ccall` shim (julia/GeoRDDB200.jl): same symbols, same
argument order.  Arrays are passed as Fortran-ordered float64 NumPy arrays (Julia's layout).
There is no CPU fallback: if the library or a B200-class GPU is missing every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgeordd_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ll_p = C.POINTER(C.c_longlong)
c_void_pp = C.POINTER(C.c_void_p)


class Kernel(C.Structure):
    """geordd_kernel POD: kind 0 SE, 1 Mat12, 2 Mat32, 3 Mat52."""
    _fields_ = [("kind", C.c_int), ("ell", C.c_double), ("sigma2", C.c_double), ("const_sigma2", C.c_double),
                ("mean_const", C.c_double)]


class GeorddError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"geordd_b200 error {code}: {msg}")
        self.code = code


class PosDefException(ArithmeticError):
    """Mirrors LinearAlgebra.PosDefException(info)."""

    def __init__(self, info, msg=""):
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed (info={info}) {msg}")
        self.info = info


# name -> (restype, argtypes)
_SIGS = {
    "geordd_version": (C.c_int, []),
    "geordd_launch_count": (C.c_longlong, []),
    "geordd_ctx_create": (C.c_int, [c_void_pp, C.c_int]),
    "geordd_ctx_destroy": (C.c_int, [C.c_void_p]),
    "geordd_last_error": (C.c_char_p, [C.c_void_p]),
    "geordd_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "geordd_ctx_sync": (C.c_int, [C.c_void_p]),
    "geordd_ctx_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "geordd_ctx_profile_get": (C.c_int, [C.c_void_p, C.c_char_p, c_double_p, c_ll_p]),
    "geordd_ctx_set_max_batch": (C.c_int, [C.c_void_p, C.c_long]),
    "geordd_ctx_set_mll_forward_only": (C.c_int, [C.c_void_p, C.c_int]),
    "geordd_ctx_last_min_margin": (C.c_int, [C.c_void_p, c_double_p]),
    "geordd_cov_sym": (C.c_int, [C.c_void_p, C.POINTER(Kernel), c_double_p, C.c_int, C.c_int, C.c_double, c_double_p]),
    "geordd_cov_cross": (C.c_int, [C.c_void_p, C.POINTER(Kernel), c_double_p, C.c_int, c_double_p, C.c_int, C.c_int,
                                   c_double_p]),
    "geordd_gp_fit": (C.c_int, [C.c_void_p, C.POINTER(Kernel), c_double_p, C.c_int, C.c_int, c_double_p, C.c_double,
                                c_void_pp, c_double_p, c_double_p, c_double_p, c_int_p]),
    "geordd_gp_fit_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Kernel), C.POINTER(c_double_p), C.c_int, c_int_p,
                                      C.POINTER(c_double_p), c_double_p, c_void_pp, C.POINTER(c_double_p), c_double_p, c_int_p,
                                      c_int_p]),
    "geordd_gp_set_y": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "geordd_gp_predict_mu": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p]),
    "geordd_gp_get_K": (C.c_int, [C.c_void_p, c_double_p]),
    "geordd_gp_nobs": (C.c_int, [C.c_void_p]),
    "geordd_gp_destroy": (C.c_int, [C.c_void_p]),
    "geordd_cliff_face": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int, c_double_p, c_double_p]),
    "geordd_inverse_variance": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, C.c_double, c_double_p,
                                          c_double_p]),
    "geordd_solve_general": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_int]),
    "geordd_weight_at_units": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p, c_double_p]),
    "geordd_projection_points": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_int, C.POINTER(C.c_int), C.c_int,
                                           c_double_p, c_double_p]),
    "geordd_points_in_region": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_int, C.POINTER(C.c_int), C.c_int,
                                          C.POINTER(C.c_ubyte)]),
    "geordd_chistat_obs": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int, c_double_p]),
    "geordd_invvar_obs": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int, C.c_double, c_double_p, c_double_p,
                                    c_double_p]),
    "geordd_invvar_calib": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int, C.c_double, c_double_p]),
    "geordd_null_prepare": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int, C.c_double, c_void_pp, c_int_p,
                                      c_double_p]),
    "geordd_null_set_sentinels": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_double, c_int_p]),
    "geordd_null_chistat_obs": (C.c_int, [C.c_void_p, c_double_p]),
    "geordd_null_invvar_obs": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "geordd_null_gp": (C.c_int, [C.c_void_p, c_void_pp]),
    "geordd_gp_prior_rand": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, c_double_p]),
    "geordd_null_chi2": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, C.c_double,
                                   c_double_p, c_ll_p]),
    "geordd_null_invvar": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, C.c_double,
                                     c_double_p, c_ll_p]),
    "geordd_null_invvar_calib": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, c_double_p]),
    "geordd_null_mll": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, C.c_double,
                                  c_double_p, c_double_p, c_ll_p, c_double_p, c_double_p]),
    "geordd_null_all": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, C.c_double,
                                  c_double_p, c_double_p, c_double_p, c_double_p]),
    "geordd_power": (C.c_int, [C.c_void_p, C.c_double, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, c_double_p,
                               c_double_p, c_double_p, C.c_long, C.c_int, c_double_p]),
    "geordd_null_destroy": (C.c_int, [C.c_void_p]),
    "geordd_multigp_create": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_int, c_int_p, C.c_int,
                                        c_double_p, c_void_pp]),
    "geordd_multigp_mll": (C.c_int, [C.c_void_p, C.POINTER(Kernel), C.c_double, C.c_double, c_double_p, c_double_p,
                                     c_int_p]),
    "geordd_multigp_mll_grad": (C.c_int, [C.c_void_p, C.POINTER(Kernel), C.c_double, C.c_double, C.c_int, C.c_int,
                                          C.c_int, C.c_int, c_double_p, c_double_p, c_int_p]),
    "geordd_multigp_postmean_beta": (C.c_int, [C.c_void_p, C.POINTER(Kernel), C.c_double, C.c_double, c_double_p]),
    "geordd_multigp_destroy": (C.c_int, [C.c_void_p]),
    "geordd_group_create": (C.c_int, [c_void_pp, C.c_int, c_int_p]),
    "geordd_group_destroy": (C.c_int, [C.c_void_p]),
    "geordd_group_size": (C.c_int, [C.c_void_p]),
    "geordd_group_last_error": (C.c_char_p, [C.c_void_p]),
    "geordd_group_ctx": (C.c_void_p, [C.c_void_p, C.c_int]),
    "geordd_group_gp_fit": (C.c_int, [C.c_void_p, C.POINTER(Kernel), c_double_p, C.c_int, C.c_int, c_double_p, C.c_double,
                                      c_void_pp, c_double_p, c_double_p, c_int_p]),
    "geordd_group_gp_destroy": (C.c_int, [C.c_void_p]),
    "geordd_group_gp_replica": (C.c_void_p, [C.c_void_p, C.c_int]),
    "geordd_group_null_prepare": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int, C.c_double, c_void_pp, c_int_p,
                                            c_double_p]),
    "geordd_group_null_destroy": (C.c_int, [C.c_void_p]),
    "geordd_group_null_replica": (C.c_void_p, [C.c_void_p, C.c_int]),
    "geordd_group_null_mll": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, C.c_double,
                                        c_double_p, c_double_p, c_ll_p]),
    "geordd_group_null_chi2": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, C.c_double,
                                         c_double_p, c_ll_p]),
    "geordd_group_null_invvar": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, C.c_double,
                                           c_double_p, c_ll_p]),
    "geordd_group_power": (C.c_int, [C.c_void_p, C.c_double, C.c_long, C.c_ulonglong, C.c_ulonglong, c_double_p, c_double_p,
                                     c_double_p, c_double_p, C.c_long, C.c_int, c_double_p]),
    "geordd_comm_unique_id": (C.c_int, [C.c_char_p]),
    "geordd_comm_init_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "geordd_comm_allreduce_tally": (C.c_int, [C.c_void_p, c_ll_p, C.c_int]),
    "geordd_comm_destroy": (C.c_int, [C.c_void_p]),
    "geordd_dbg_guard_violations": (C.c_longlong, []),
    "geordd_dbg_set_lu_blocked": (C.c_int, [C.c_int]),
    "geordd_dbg_normals": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_ulonglong, C.c_int, C.c_int, c_double_p]),
    "geordd_dbg_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, c_double_p,
                                  C.c_int, c_double_p, C.c_int, C.c_double, c_double_p, C.c_int, C.c_int]),
    "geordd_dbg_potrf": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p, c_double_p]),
    "geordd_dbg_trsm": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p]),
    "geordd_bench_potrf": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    "geordd_bench_tileops": (C.c_int, [C.c_void_p, c_double_p]),
    "geordd_bench_dmma_peak": (C.c_int, [C.c_void_p, c_double_p]),
    "geordd_bench_dmma": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS.keys())

_lib = None


def load_library(path: str = None):
    """dlopen the library (fails loudly if it was not built) and attach prototypes.  GEORDD_B200_LIB selects
    another build of the same library (kernel experiments)."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("GEORDD_B200_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise GeorddError(-2, f"{path} not found: build it with `python geordd.jl_b200/build.py` "
                              "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    """Pointer to a float64 array (None -> NULL).  The array must be F-contiguous (or 1-d)."""
    if a is None:
        return None
    assert a.dtype == np.float64
    assert a.flags.f_contiguous or a.flags.c_contiguous and a.ndim == 1, "array must be column-major"
    return a.ctypes.data_as(c_double_p)


def fmat(a):
    """Column-major float64 copy/view."""
    return np.asfortranarray(a, dtype=np.float64)
