"""ctypes binding of libgpsat_b200.so (the C ABI declared in include/gpsat_b200.h).

The library is built in-tree by gpsat_b200/build.py (nvcc, sm_100a).  There is NO CPU fallback:
if the shared object is missing or cannot be loaded this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpsat_b200.so")

MAXD = 4
MAXP = 6
SEL_MAXTERMS = 8

KERNEL_IDS = {"Matern32": 0, "Matern52": 1, "Matern12": 2, "Exponential": 2, "RBF": 3,
              "SquaredExponential": 3}
OPT_STATUS = {0: "RUNNING", 1: "CONVERGENCE: NORM_OF_PROJECTED_GRADIENT_<=_PGTOL",
              2: "CONVERGENCE: REL_REDUCTION_OF_F_<=_FACTR*EPSMCH",
              3: "STOP: TOTAL NO. of ITERATIONS REACHED LIMIT",
              4: "STOP: TOTAL NO. of f AND g EVALUATIONS EXCEEDS LIMIT",
              5: "ABNORMAL_TERMINATION_IN_LNSRCH"}


class Batch(C.Structure):
    _fields_ = [("n_experts", C.c_int), ("D", C.c_int), ("kernel_id", C.c_int), ("obs_mean_local", C.c_int),
                ("offsets_host", C.c_void_p), ("offsets_dev", C.c_void_p), ("coords_dev", C.c_void_p),
                ("obs_dev", C.c_void_p), ("coords_scale", C.c_double * MAXD), ("obs_scale", C.c_double),
                ("obs_mean_out_dev", C.c_void_p)]


class CellGrid(C.Structure):
    _fields_ = [("x0", C.c_double), ("y0", C.c_double), ("cell", C.c_double), ("ncx", C.c_int), ("ncy", C.c_int)]


class SgprBatch(C.Structure):
    _fields_ = [("data", Batch), ("z_offsets_host", C.c_void_p), ("z_offsets_dev", C.c_void_p),
                ("z_coords_dev", C.c_void_p)]


class Transforms(C.Structure):
    _fields_ = [("kind", C.c_int * MAXP), ("low", C.c_double * MAXP), ("high", C.c_double * MAXP),
                ("trainable", C.c_int * MAXP)]


class OptOptions(C.Structure):
    _fields_ = [("maxcor", C.c_int), ("maxiter", C.c_int), ("maxfun", C.c_int), ("maxls", C.c_int),
                ("ftol", C.c_double), ("gtol", C.c_double)]


class SelTerm(C.Structure):
    _fields_ = [("type", C.c_int), ("ncol", C.c_int), ("comp", C.c_int), ("pad_", C.c_int),
                ("col", C.c_int * 4), ("rcol", C.c_int * 4), ("val", C.c_double)]


class SelSpec(C.Structure):
    _fields_ = [("nterms", C.c_int), ("pad_", C.c_int), ("t", SelTerm * SEL_MAXTERMS)]


_EXPORTS = {
    "gpsat_last_error": (C.c_char_p, []),
    "gpsat_version": (C.c_int, []),
    "gpsat_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_size_t]),
    "gpsat_destroy": (C.c_int, [C.c_void_p]),
    "gpsat_default_opts": (None, [C.POINTER(OptOptions)]),
    "gpsat_select_count": (C.c_int, [C.POINTER(SelSpec), C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "gpsat_select_fill": (C.c_int, [C.POINTER(SelSpec), C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpsat_bucket_build": (C.c_int, [C.POINTER(SelSpec), C.POINTER(CellGrid), C.c_void_p, C.c_longlong, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpsat_select_bucket": (C.c_int, [C.POINTER(SelSpec), C.POINTER(CellGrid), C.c_void_p, C.c_longlong, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "gpsat_gather_rows": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                    C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpsat_gather_pred": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_void_p]),
    "gpsat_kernel_matrix": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_void_p]),
    "gpsat_gpr_eval": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpsat_gpr_optimise": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(Transforms),
                                     C.POINTER(OptOptions), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "gpsat_gpr_predict": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpsat_gpr_predict_cov": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "gpsat_sgpr_eval": (C.c_int, [C.c_void_p, C.POINTER(SgprBatch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpsat_sgpr_optimise": (C.c_int, [C.c_void_p, C.POINTER(SgprBatch), C.c_void_p, C.POINTER(Transforms),
                                      C.POINTER(OptOptions), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "gpsat_sgpr_predict": (C.c_int, [C.c_void_p, C.POINTER(SgprBatch), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpsat_debug_factor": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "gpsat_launch_count": (C.c_longlong, [C.c_void_p]),
    "gpsat_sync_timeouts": (C.c_longlong, [C.c_void_p]),
    "gpsat_last_plan": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_size_t)]),
    "gpsat_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "gpsat_get_profile": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_double)] * 9),
    "gpsat_gaussian_smooth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.c_void_p, C.c_void_p]),
    "gpsat_weighted_groups": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_void_p, C.c_void_p]),
    "gpsat_bin_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int,
                                       C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpsat_bin_spread": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int,
                                   C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gpsat_dmma_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "gpsat_microbench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "gpsat_lbfgs_state_bytes": (C.c_size_t, []),
    "gpsat_lbfgs_init_host": (None, [C.c_void_p, C.c_void_p, C.c_int]),
    "gpsat_lbfgs_tell_host": (C.c_int, [C.c_void_p, C.POINTER(OptOptions), C.c_double, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}

_lib = None


def exported_symbols():
    return sorted(_EXPORTS)


def load():
    """Load the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m gpsat_b200.build` "
            "(gpsat_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class GpsatError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().gpsat_last_error()
        raise GpsatError(f"gpsat_b200 error {rc}: {msg.decode() if msg else ''}")
