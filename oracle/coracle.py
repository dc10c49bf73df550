"""
ORACLE (test infrastructure) -- ctypes front end of oracle/tri_oracle.c (the plain-C restatement
of likelihoods.py:302-587 over the restated transit model).  PARITY UNPINNED, see
oracle/quadmodel.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtri_oracle.so")
_lib = None

_D = ctypes.POINTER(ctypes.c_double)
_I64 = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "tri_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.tro_eval_quad.restype = ctypes.c_double
        L.tro_eval_quad.argtypes = [ctypes.c_double] * 4
        L.tro_z.restype = ctypes.c_double
        L.tro_z.argtypes = [ctypes.c_double] * 6
        L.tro_log_mean_exp.restype = ctypes.c_double
        L.tro_log_mean_exp.argtypes = [_D, ctypes.c_int64]
        L.tro_num_threads.restype = ctypes.c_int
        L.tro_set_num_threads.argtypes = [ctypes.c_int]
        L.tro_make_table.argtypes = [_D, _D, _D]
        L.tro_model.argtypes = [ctypes.c_int64, _D] + [ctypes.c_double] * 9 + [ctypes.c_int, _D]
        L.tro_lnl_tp.argtypes = ([ctypes.c_int64, _D, _D, ctypes.c_double, ctypes.c_double,
                                  ctypes.c_int, ctypes.c_int64] + [_D] * 10
                                 + [ctypes.c_int, _D, _I64])
        L.tro_lnl_eb_rule.argtypes = ([ctypes.c_int64, _D, _D, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_int, ctypes.c_int64] + [_D] * 11
                                      + [ctypes.c_int, ctypes.c_int, ctypes.c_int, _D, _D, _I64])
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_D)


def _arr(x, n=None):
    a = np.ascontiguousarray(x, dtype=np.float64)
    if n is not None and a.ndim == 0:
        a = np.full(n, float(a))
    return a


def orbit_table():
    es = np.empty(256)
    ms = np.empty(512)
    tae = np.empty((256, 512))
    lib().tro_make_table(_p(es), _p(ms), _p(tae))
    return es, ms, tae


def eval_quad(z, k, u1, u2):
    return lib().tro_eval_quad(z, k, u1, u2)


def model(time, k, p, a_rs, inc_rad, e, w_rad, u1, u2, exptime=0.0, nsamples=1):
    time = _arr(time)
    out = np.empty_like(time)
    lib().tro_model(time.size, _p(time), k, p, a_rs, inc_rad, e, w_rad, u1, u2, exptime,
                    int(nsamples), _p(out))
    return out


def lnL_TP_p(time, flux, sigma, R_p, P_orb, inc, a, R_s, u1, u2, ecc, argp,
             companion_fluxratio, companion_is_host=False, exptime=0.00139, nsamples=20,
             counts=None):
    """Same signature and return (+0.5 chi^2) as the reference's lnL_TP_p (likelihoods.py:443)."""
    time, flux = _arr(time), _arr(flux)
    n = np.size(R_p)
    A = [_arr(x, n) for x in (R_p, P_orb, inc, a, R_s, u1, u2, ecc, argp, companion_fluxratio)]
    out = np.empty(n)
    cp = counts.ctypes.data_as(_I64) if counts is not None else None
    lib().tro_lnl_tp(time.size, _p(time), _p(flux), float(sigma), float(exptime), int(nsamples),
                     n, *[_p(x) for x in A], int(bool(companion_is_host)), _p(out), cp)
    return out


def _lnl_eb(twin, time, flux, sigma, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp,
            companion_fluxratio, companion_is_host, exptime, nsamples, counts, secdepth,
            scalar_rule=False):
    time, flux = _arr(time), _arr(flux)
    n = np.size(R_EB)
    A = [_arr(x, n) for x in (R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp,
                              companion_fluxratio)]
    out = np.empty(n)
    cp = counts.ctypes.data_as(_I64) if counts is not None else None
    sp = _p(secdepth) if secdepth is not None else None
    lib().tro_lnl_eb_rule(time.size, _p(time), _p(flux), float(sigma), float(exptime),
                          int(nsamples), n, *[_p(x) for x in A], int(bool(companion_is_host)),
                          int(twin), int(bool(scalar_rule)), _p(out), sp, cp)
    return out


def lnL_EB_p(time, flux, sigma, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp,
             companion_fluxratio, companion_is_host=False, exptime=0.00139, nsamples=20,
             counts=None, secdepth=None, scalar_rule=False):
    """Reference lnL_EB_p (likelihoods.py:490): +0.5 chi^2, +inf where secdepth >= 1.5 sigma.
    scalar_rule: the radius-ratio rules of the scalar lnL_EB (likelihoods.py:121-123, :137)."""
    return _lnl_eb(0, time, flux, sigma, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc,
                   argp, companion_fluxratio, companion_is_host, exptime, nsamples, counts,
                   secdepth, scalar_rule)


def lnL_EB_twin_p(time, flux, sigma, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp,
                  companion_fluxratio, companion_is_host=False, exptime=0.00139, nsamples=20,
                  counts=None, secdepth=None, scalar_rule=False):
    """Reference lnL_EB_twin_p (likelihoods.py:542)."""
    return _lnl_eb(1, time, flux, sigma, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc,
                   argp, companion_fluxratio, companion_is_host, exptime, nsamples, counts,
                   secdepth, scalar_rule)


def log_mean_exp(logw):
    logw = _arr(logw)
    return lib().tro_log_mean_exp(_p(logw), logw.size)


def num_threads():
    return lib().tro_num_threads()


def set_num_threads(n):
    """Threads of the OpenMP loops over draws (0 = the process-wide OpenMP setting)."""
    lib().tro_set_num_threads(int(n))
