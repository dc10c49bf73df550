"""ctypes binding of include/triceratops_b200.h (the drop-in boundary of the engine).

There is no CPU fallback: loading fails loudly if the CUDA library has not been built, and
every call raises if no GPU is usable.
"""
import ctypes
import os

import numpy as np

from . import _build

c_double_p = ctypes.POINTER(ctypes.c_double)
c_uint8_p = ctypes.POINTER(ctypes.c_uint8)

TRI_OK, TRI_ECUDA, TRI_EINVAL, TRI_ESTATE, TRI_ENODEVICE = 0, -1, -2, -3, -4
TRI_MAX_INFLIGHT = 4     # include/triceratops_b200.h


class TriError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("triceratops_b200 error %d: %s" % (code, msg))
        self.code = code


class tri_col(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("stride", ctypes.c_int64)]


class tri_tp_args(ctypes.Structure):
    _fields_ = ([("N", ctypes.c_int64)]
                + [(n, tri_col) for n in ("rp", "P_orb", "inc", "ecc", "argp", "mtot", "rhost",
                                          "u1", "u2", "cfr", "lnprior")]
                + [("extra_mask", ctypes.c_void_p), ("companion_is_host", ctypes.c_int32)])


class tri_eb_args(ctypes.Structure):
    _fields_ = ([("N", ctypes.c_int64)]
                + [(n, tri_col) for n in ("reb", "ebfr", "q", "P_orb", "inc", "ecc", "argp",
                                          "mtot", "rhost", "u1", "u2", "cfr", "lnprior")]
                + [("extra_mask", ctypes.c_void_p), ("companion_is_host", ctypes.c_int32),
                   ("scalar_loop", ctypes.c_int32)])


class tri_result(ctypes.Structure):
    _fields_ = [("lnZ", ctypes.c_double), ("m", ctypes.c_double), ("s", ctypes.c_double),
                ("n_finite", ctypes.c_int64), ("n_posinf", ctypes.c_int64),
                ("n_pass", ctypes.c_int64), ("n_stamps", ctypes.c_int64),
                ("n_interior", ctypes.c_int64), ("n_limb", ctypes.c_int64),
                ("lnL_out", ctypes.c_void_p), ("mask_out", ctypes.c_void_p),
                ("top_cap", ctypes.c_int64), ("n_top", ctypes.c_int64),
                ("n_evaluated", ctypes.c_int64),
                ("top_idx", ctypes.c_void_p), ("top_lnL", ctypes.c_void_p)]


# every symbol include/triceratops_b200.h declares
EXPORTS = ("tri_init", "tri_shutdown", "tri_last_error", "tri_set_lightcurve", "tri_eval_tp",
           "tri_eval_eb", "tri_eval_tp_dev", "tri_eval_eb_dev", "tri_lnl_tp", "tri_lnl_eb",
           "tri_simulate_tp", "tri_simulate_eb",
           "tri_fetch_lnl", "tri_log_mean_exp", "tri_last_timing", "tri_fp64_peak", "tri_sm_count",
           "tri_submit_tp", "tri_submit_eb", "tri_submit_tp_dev", "tri_submit_eb_dev", "tri_wait",
           "tri_dev_splev", "tri_set_counting", "tri_set_lightcurve_err")

_lib = None


def library_path():
    # TRI_B200_LIB lets developers A/B a differently tuned build of the same sources
    return os.environ.get("TRI_B200_LIB", _build.SO_PATH)


def load():
    """Load the CUDA library (raises if it is missing -- build it with __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            "CUDA extension %s is missing; run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc). triceratops_b200 has no CPU fallback." % path)
    L = ctypes.CDLL(path)
    for name in EXPORTS:
        getattr(L, name)  # AttributeError if the header and the library disagree
    L.tri_last_error.restype = ctypes.c_char_p
    L.tri_init.argtypes = [ctypes.c_int]
    L.tri_set_lightcurve.argtypes = [c_double_p, c_double_p, ctypes.c_int64, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_int32]
    L.tri_set_lightcurve_err.argtypes = [c_double_p, c_double_p, c_double_p, ctypes.c_int64,
                                         ctypes.c_double, ctypes.c_int32]
    L.tri_eval_tp.argtypes = [ctypes.POINTER(tri_tp_args), ctypes.POINTER(tri_result)]
    L.tri_eval_eb.argtypes = [ctypes.POINTER(tri_eb_args), ctypes.POINTER(tri_result)]
    L.tri_eval_tp_dev.argtypes = [ctypes.POINTER(tri_tp_args), ctypes.POINTER(tri_result),
                                  ctypes.c_void_p]
    L.tri_eval_eb_dev.argtypes = [ctypes.POINTER(tri_eb_args), ctypes.POINTER(tri_result),
                                  ctypes.c_void_p]
    c_i64_p = ctypes.POINTER(ctypes.c_int64)
    L.tri_submit_tp.argtypes = [ctypes.POINTER(tri_tp_args), ctypes.POINTER(tri_result), c_i64_p]
    L.tri_submit_eb.argtypes = [ctypes.POINTER(tri_eb_args), ctypes.POINTER(tri_result), c_i64_p]
    L.tri_submit_tp_dev.argtypes = [ctypes.POINTER(tri_tp_args), ctypes.POINTER(tri_result),
                                    ctypes.c_void_p, c_i64_p]
    L.tri_submit_eb_dev.argtypes = [ctypes.POINTER(tri_eb_args), ctypes.POINTER(tri_result),
                                    ctypes.c_void_p, c_i64_p]
    L.tri_wait.argtypes = [ctypes.c_int64, ctypes.POINTER(tri_result)]
    L.tri_dev_splev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    L.tri_lnl_tp.argtypes = [ctypes.c_int64] + [c_double_p] * 10 + [ctypes.c_int32, c_double_p]
    L.tri_lnl_eb.argtypes = ([ctypes.c_int64] + [c_double_p] * 11
                             + [ctypes.c_int32, ctypes.c_int32, c_double_p])
    L.tri_simulate_tp.argtypes = [ctypes.c_int64] + [c_double_p] * 10 + [ctypes.c_int32,
                                                                        c_double_p]
    L.tri_simulate_eb.argtypes = ([ctypes.c_int64] + [c_double_p] * 11
                                  + [ctypes.c_int32, ctypes.c_int32, c_double_p, c_double_p])
    L.tri_fetch_lnl.argtypes = [ctypes.c_int32, c_double_p, ctypes.c_int64]
    L.tri_log_mean_exp.argtypes = [c_double_p, ctypes.c_int64, ctypes.POINTER(tri_result)]
    L.tri_last_timing.argtypes = [c_double_p, c_double_p, c_double_p,
                                  ctypes.POINTER(ctypes.c_int32)]
    L.tri_fp64_peak.argtypes = [c_double_p]
    L.tri_set_counting.argtypes = [ctypes.c_int32]
    L.tri_sm_count.argtypes = [ctypes.POINTER(ctypes.c_int32)]
    _lib = L
    return L


def check(rc):
    if rc != TRI_OK:
        raise TriError(rc, load().tri_last_error().decode("utf-8", "replace"))


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def f64(x):
    """C-contiguous float64 view/copy (what the ABI requires of every column)."""
    return np.ascontiguousarray(x, dtype=np.float64)
