"""A scenario's element-wise preparation in one GIL-free call (csrc/host_blocks.c).

The lnZ_* functions of marginal_likelihoods.py turn their prior deviates into engine columns
with ~200 numpy statements per scenario.  Run chunk by chunk in Python threads those statements
contend for the GIL (16 workers spend two thirds of their time waiting for it); here the same
statements run in C, in the same order, over OpenMP threads.  The arithmetic is IEEE and
compiled without contraction; the transcendental functions are numpy's own compiled inner
loops, looked up in the ufunc objects (numpy.power, log10, log, exp, arccos) and called
directly, so every value has the bits the numpy path produces.

`available()` checks that claim once per process on a small instance of every scenario kind
against the numpy path; if the library is missing, numpy's ufunc layout is not the expected
one, or any bit differs, `run()` returns None and the callers use the numpy path.  This is
host-side preparation of the prior draws, not the light-curve path.
"""
import ctypes
import os
import threading

import numpy as np

from . import _build
from ._bufpool import empty as _pinned_empty

KINDS = ("TTP", "TEB", "PTP", "PEB", "STP", "SEB", "DTP", "DEB", "BTP", "BEB")
#: number of float64 output columns per kind, and whether a boolean `extra` mask comes with them
N_OUT = {"TTP": 3, "TEB": 7, "PTP": 5, "PEB": 9, "STP": 9, "SEB": 13, "DTP": 5, "DEB": 9,
         "BTP": 9, "BEB": 13}
HAS_EXTRA = {"PTP", "PEB", "STP", "SEB", "BTP", "BEB"}
MIN_N = 32768

_D = ctypes.POINTER(ctypes.c_double)
_I64 = ctypes.POINTER(ctypes.c_int64)


class _NpFn(ctypes.Structure):
    _fields_ = [("fn", ctypes.c_void_p), ("data", ctypes.c_void_p)]


class _Spline(ctypes.Structure):
    _fields_ = [("t", _D), ("c", _D), ("n", ctypes.c_int32), ("k", ctypes.c_int32)]


class _Powerlaw(ctypes.Structure):
    _fields_ = [("nseg", ctypes.c_int32), ("pad", ctypes.c_int32), ("fill", ctypes.c_double),
                ("norm", ctypes.c_double)] + [
        (name, ctypes.c_double * 3) for name in ("knot", "integ", "p1", "amp", "e0", "inv")]


class _Bound(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("m_ge_1", ctypes.c_int32), ("d", ctypes.c_double),
                ("xp", _D), ("fp", _D), ("nxp", ctypes.c_int64)] + [
        (name, ctypes.c_double) for name in (
            "K1", "au", "f1", "f2", "f3", "slope", "slope2", "two_f1", "half_alpha",
            "alpha_dlogP", "t2", "t23", "t234", "t2345", "t4", "t45", "M_act")]


class _BackgroundPrior(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("pad", ctypes.c_int32), ("constant", ctypes.c_double),
                ("K", ctypes.c_double), ("xp", _D), ("fp", _D), ("nxp", ctypes.c_int64)]


class _Ldc(ctypes.Structure):
    _fields_ = [("code", _I64), ("u1", _D), ("u2", _D), ("n", ctypes.c_int64),
                ("cap", ctypes.c_double)]


class _Args(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("flatpriors", ctypes.c_int32),
                ("has_cc", ctypes.c_int32), ("molusc", ctypes.c_int32), ("N", ctypes.c_int64),
                ("nthreads", ctypes.c_int32), ("pad", ctypes.c_int32)] + [
        (name, _NpFn) for name in ("f_pow", "f_log10", "f_log", "f_exp", "f_arccos")] + [
        (name, _D) for name in ("x_rp", "x_inc", "x_q", "x_w", "c_comp", "x_e")] + [
        ("idxs", _I64)] + [
        (name, ctypes.c_double) for name in ("M_s", "R_s", "Teff", "c_lo", "inc_norm", "ecc_exp",
                                             "G", "Msun", "Rsun", "rp_flat_A")] + [
        (name, _Powerlaw) for name in ("rp_hi", "rp_lo", "q", "q_comp")] + [
        (name, _Spline) for name in ("hot_R", "hot_T", "cool_R", "cool_T", "flux_tess",
                                     "flux_cc")] + [
        ("f0_tess", ctypes.c_double), ("f0_cc", ctypes.c_double), ("bound", _Bound),
        ("bgp", _BackgroundPrior), ("ldc", _Ldc), ("ntab", ctypes.c_int64)] + [
        (name, _D) for name in ("bg_mass", "bg_radius", "bg_teff", "bg_logg", "bg_fr", "bg_band",
                                "bg_fr_tess", "bg_fr_cc", "bg_u1", "bg_u2")] + [
        ("out", _D * 16), ("extra", ctypes.POINTER(ctypes.c_uint8)), ("interp_delta", _D),
        ("interp_j", ctypes.POINTER(ctypes.c_int32))]


# ---- numpy's compiled inner loops ------------------------------------------------------------
_NPY_DOUBLE = 12


def _double_loop(ufunc):
    """(function pointer, data pointer) of the first all-float64 loop registered with `ufunc`
    (the one its type resolver picks for float64 operands), read from the PyUFuncObject:
    PyObject_HEAD | int nin, nout, nargs, identity | functions | data | int ntypes, reserved |
    name | types (numpy/ufuncobject.h, unchanged since numpy 1.x)."""
    base = id(ufunc)
    nin, nout, nargs, _ = (ctypes.c_int * 4).from_address(base + 16)
    if (nin, nout) != (ufunc.nin, ufunc.nout) or nargs != nin + nout:
        raise RuntimeError("unexpected PyUFuncObject layout")
    functions = ctypes.c_void_p.from_address(base + 32).value
    data = ctypes.c_void_p.from_address(base + 40).value
    ntypes = ctypes.c_int.from_address(base + 48).value
    name = ctypes.c_char_p.from_address(base + 56).value
    types = ctypes.c_void_p.from_address(base + 64).value
    if ntypes != ufunc.ntypes or name != ufunc.__name__.encode() or not functions or not types:
        raise RuntimeError("unexpected PyUFuncObject layout")
    sigs = (ctypes.c_char * (ntypes * nargs)).from_address(types).raw
    want = bytes([_NPY_DOUBLE]) * nargs
    for i in range(ntypes):
        if sigs[i * nargs:(i + 1) * nargs] == want:
            fn = ctypes.c_void_p.from_address(functions + 8 * i).value
            dt = ctypes.c_void_p.from_address(data + 8 * i).value if data else None
            if fn:
                return _NpFn(fn, dt)
    raise RuntimeError("no float64 loop in numpy.%s" % ufunc.__name__)


_lib = None
_loops = None
_state = None          # None: not tried; True / False afterwards


def _load():
    global _lib, _loops
    if _lib is None:
        L = ctypes.CDLL(_build.HOST_SO_PATH)
        L.trih_scenario_block.argtypes = [ctypes.POINTER(_Args)]
        L.trih_scenario_block.restype = ctypes.c_int
        L.trih_block_args_size.restype = ctypes.c_int64
        if L.trih_block_args_size() != ctypes.sizeof(_Args):
            raise RuntimeError("tb_args layout mismatch")
        _loops = {n: _double_loop(getattr(np, n))
                  for n in ("power", "log10", "log", "exp", "arccos")}
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(_D)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _spline(sp, keep):
    t, c, k = sp._eval_args
    t, c = _f64(t), _f64(c)
    keep += [t, c]
    return _Spline(_dp(t), _dp(c), int(t.size), int(k))


def _powerlaw(spec, keep):
    """tb_powerlaw from (edges, powers, amps), with the scalars the numpy path computes
    (priors._piecewise_powerlaw); None: every element becomes 1.0."""
    from .priors import powerlaw_tables
    P = _Powerlaw()
    if spec is None:
        P.nseg, P.fill = 0, 1.0
        return P
    edges, powers, amps = spec
    integrals, cum, norm = powerlaw_tables(edges, powers, amps)
    P.nseg, P.norm = len(powers), norm
    for k in range(len(powers)):
        p1 = powers[k] + 1
        P.knot[k] = norm * cum[k]
        P.integ[k] = integrals[k]
        P.p1[k] = p1
        P.amp[k] = amps[k]
        P.e0[k] = edges[k] ** p1
        P.inv[k] = 1 / p1
    return P


def order_free(contrasts):
    """numpy.interp starts each search where the previous query ended; on a table that is not
    non-decreasing the interval found can depend on that, i.e. on the order and grouping of the
    queries.  True when it cannot (the queries may then be evaluated in independent chunks)."""
    c = np.asarray(contrasts, dtype=float)
    return bool(c.size < 2 or np.all(np.diff(c) >= 0))


class Population:
    """Background-star tables of the D*/B* scenarios as contiguous float64 arrays."""

    def __init__(self, masses, radii, teffs, loggs, fluxratios, band, fr_tess, fr_cc, u1s, u2s):
        self.n = int(len(fluxratios))
        self.cols = [None if a is None else _f64(a)
                     for a in (masses, radii, teffs, loggs, fluxratios, band, fr_tess, fr_cc,
                               u1s, u2s)]


def _block_threads(n):
    """OpenMP threads of one scenario block (TRI_B200_BLOCK_THREADS; default: the host threads)."""
    v = os.environ.get("TRI_B200_BLOCK_THREADS")
    return max(1, int(v)) if v else n


def run(kind, N, *, M_s, R_s, Teff, x_inc, x_w, x_rp=None, x_q=None, x_e=None, c_comp=None,
        idxs=None, flatpriors=False, molusc=False, P_mean=None, filt="TESS",
        contrast_curve_file=None, plx=None, bound_kind=None, ldc_grid=None, Z=None,
        ldc_cap=None, population=None, N_comp=None, _force=False):
    """Outputs of scenario `kind`'s block for N draws: tuple of float64 columns (+ the boolean
    extra mask last, for the kinds that have one) in the order of the block functions of
    marginal_likelihoods.py -- or None when the C path is not usable for this call."""
    if not _force and (N < MIN_N or not available()):
        return None
    from ._constants import G, Msun, Rsun, au, pi
    from ._hostpar import N_THREADS, _inline
    from . import funcs, priors
    L = _load()
    keep = []
    A = _Args()
    A.kind, A.N = KINDS.index(kind), int(N)
    A.nthreads = 1 if _inline() else _block_threads(int(N_THREADS))
    A.flatpriors = int(bool(flatpriors))
    A.molusc = int(bool(molusc))
    A.has_cc = int(contrast_curve_file is not None)
    for name in ("power", "log10", "log", "exp", "arccos"):
        setattr(A, "f_" + {"power": "pow"}.get(name, name), _loops[name])

    def col(a, dtype=np.float64):
        if a is None:
            return None
        if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.ndim == 1
                and a.shape[0] == N and a.flags.c_contiguous):
            raise TypeError("block column")
        keep.append(a)
        return a

    try:
        for name, a in (("x_rp", x_rp), ("x_inc", x_inc), ("x_q", x_q), ("x_w", x_w),
                        ("c_comp", c_comp), ("x_e", x_e)):
            a = col(a)
            if a is not None:
                setattr(A, name, _dp(a))
        if idxs is not None:
            A.idxs = col(idxs, np.int64).ctypes.data_as(_I64)
    except TypeError:
        return None
    if flatpriors not in (True, False):
        return None
    A.M_s, A.R_s, A.Teff = float(M_s), float(R_s), float(Teff)
    A.G, A.Msun, A.Rsun = G, Msun, Rsun
    # sample_inc's scalars (priors.sample_inc with lower = 0, upper = 90)
    c_lo = np.cos(0 * np.pi / 180)
    A.c_lo, A.inc_norm = c_lo, 1 / (c_lo - np.cos(90 * np.pi / 180))
    if P_mean is not None:
        A.ecc_exp = 1.0 / (0.2 if P_mean <= 10 else 0.6)
    A.rp_flat_A = priors.RP_FLAT_A
    hi, lo = priors.rp_specs()
    A.rp_hi, A.rp_lo = _powerlaw(hi, keep), _powerlaw(lo, keep)
    A.q = _powerlaw(priors.mass_ratio_spec(M_s, *priors.Q_LAW), keep)
    A.q_comp = _powerlaw(priors.mass_ratio_spec(M_s, *priors.Q_COMPANION_LAW), keep)
    A.hot_R, A.hot_T = _spline(funcs._hot_R, keep), _spline(funcs._hot_T, keep)
    A.cool_R, A.cool_T = _spline(funcs._cool_R, keep), _spline(funcs._cool_T, keep)
    cc_band = filt
    if kind == "BEB":
        cc_band = filt if filt in ("J", "H", "K") else "TESS"
    if cc_band not in funcs._FLUX_SPLINES:
        return None
    A.flux_tess = _spline(funcs._FLUX_SPLINES["TESS"], keep)
    A.flux_cc = _spline(funcs._FLUX_SPLINES[cc_band], keep)
    A.f0_tess = float(funcs.flux_relation(np.array([M_s]), "TESS")[0])
    A.f0_cc = float(funcs.flux_relation(np.array([M_s]), cc_band)[0])

    if kind in ("PTP", "PEB", "STP", "SEB"):
        B = A.bound
        if molusc:
            B.mode = 0
        else:
            K = priors.bound_constants(M_s, plx)
            B.mode = 2 if bound_kind == "EB" else 1
            B.m_ge_1 = int(bool(K["M_act"] >= 1.0))
            B.d, B.M_act, B.au = K["d"], K["M_act"], au
            if contrast_curve_file is None:
                seps, cons = np.array([2.2]), np.array([1.0])
            else:
                seps, cons = funcs.file_to_contrast_curve(contrast_curve_file)
            seps, cons = _f64(seps), _f64(cons)
            keep += [seps, cons]
            B.xp, B.fp, B.nxp = _dp(cons), _dp(seps), int(cons.size)
            B.K1 = (4 * pi ** 2) / (G * K["M_s"] * Msun)
            f1, f2, f3, alpha, dlogP = K["f1"], K["f2"], K["f3"], K["alpha"], K["dlogP"]
            t2, t3, t4, t5 = K["t2"], K["t3"], K["t4"], K["t5"]
            B.f1, B.f2, B.f3, B.slope, B.slope2 = f1, f2, f3, K["slope"], K["slope2"]
            B.two_f1, B.half_alpha, B.alpha_dlogP = 2.0 * f1, 0.5 * alpha, alpha * dlogP
            B.t2, B.t23, B.t234, B.t2345 = t2, t2 + t3, t2 + t3 + t4, t2 + t3 + t4 + t5
            B.t4, B.t45 = t4, t4 + t5
    if kind in ("STP", "SEB"):
        try:
            codes, u1, u2 = ldc_grid.node_table(Z)
        except ValueError:
            return None
        keep += [codes, u1, u2]
        A.ldc = _Ldc(codes.ctypes.data_as(_I64), _dp(u1), _dp(u2), int(codes.size),
                     float(ldc_cap))
    if kind in ("DTP", "DEB", "BTP", "BEB"):
        pop = population
        keep.append(pop)
        A.ntab = pop.n
        for name, a in zip(("bg_mass", "bg_radius", "bg_teff", "bg_logg", "bg_fr", "bg_band",
                            "bg_fr_tess", "bg_fr_cc", "bg_u1", "bg_u2"), pop.cols):
            if a is not None:
                setattr(A, name, _dp(a))
        Bg = A.bgp
        if contrast_curve_file is None:
            Bg.mode = 0
            Bg.constant = np.log((N_comp / 0.1) * (1 / 3600) ** 2 * 2.2 ** 2)
        else:
            seps, cons = funcs.file_to_contrast_curve(contrast_curve_file)
            seps, cons = _f64(seps), _f64(cons)
            keep += [seps, cons]
            Bg.mode, Bg.K = 1, (N_comp / 0.1) * (1 / 3600) ** 2
            Bg.xp, Bg.fp, Bg.nxp = _dp(cons), _dp(seps), int(cons.size)

    # a contrast table numpy.interp cannot search order-free: the C side records its searches
    # and re-walks the chunk boundaries afterwards, as one whole-array call would have
    table = None
    if kind in ("PTP", "PEB", "STP", "SEB") and A.bound.mode != 0:
        table = cons
    elif kind in ("DTP", "DEB", "BTP", "BEB") and A.bgp.mode == 1:
        table = cons
    if table is not None and not order_free(table):
        delta, jrec = np.empty(N), np.empty(N, dtype=np.int32)
        keep += [delta, jrec]
        A.interp_delta, A.interp_j = _dp(delta), jrec.ctypes.data_as(
            ctypes.POINTER(ctypes.c_int32))
    # (page-locked when possible: the engine then uploads them without a staging copy)
    outs = [_pinned_empty(N) for _ in range(N_OUT[kind])]
    for i, o in enumerate(outs):
        A.out[i] = _dp(o)
    extra = None
    if kind in HAS_EXTRA:
        extra = np.empty(N, dtype=bool)
        A.extra = extra.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    rc = L.trih_scenario_block(ctypes.byref(A))
    del keep
    if rc != 0:
        return None          # (the numpy path raises what the reference raises)
    return tuple(outs) + ((extra,) if extra is not None else ())


# ---- once per process: the C path against the numpy path ---------------------------------------
_state_lock = threading.Lock()


def available():
    """True when run() may be used (checked once per process; scenario threads that arrive
    while the check runs wait for its verdict)."""
    global _state
    if _state is None:
        with _state_lock:
            if _state is None:
                ok = False
                if (os.environ.get("TRI_B200_NUMPY_BLOCKS") != "1"
                        and os.path.exists(_build.HOST_SO_PATH)):
                    try:
                        _load()
                        from . import _blocks_check
                        ok = bool(_blocks_check.self_check())
                    except Exception:
                        ok = False
                _state = ok
    return _state
