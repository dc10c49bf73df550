"""Structure of the device model (csrc/tri_model.cuh) exercised on the CPU: the transit window,
the prefix-sum treatment of out-of-transit stamps, the merged case III/IV evaluation and the
reciprocal Bulirsch sweep must reproduce the oracle's likelihoods (TEST-ONLY host build, see
tests/hostcheck/hostcheck.cpp)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import KEP10, ROOT, TOI465, load_lc
from oracle import coracle

D = ctypes.POINTER(ctypes.c_double)
I64 = ctypes.POINTER(ctypes.c_int64)


@pytest.fixture(scope="module")
def hc():
    d = os.path.join(ROOT, "tests", "hostcheck")
    subprocess.check_call(["make", "-C", d, "-s"])
    H = ctypes.CDLL(os.path.join(d, "libhostcheck.so"))
    H.hc_occult_quad.restype = ctypes.c_double
    H.hc_occult_quad.argtypes = [ctypes.c_double] * 4
    H.hc_z.restype = ctypes.c_double
    H.hc_z.argtypes = [ctypes.c_double] * 6
    H.hc_lnl.argtypes = ([ctypes.c_int, ctypes.c_int64, D, D, ctypes.c_double, ctypes.c_double,
                          ctypes.c_int, ctypes.c_int64] + [D] * 11 + [ctypes.c_int] * 3 + [D, I64])
    return H


def _p(a):
    return a.ctypes.data_as(D)


def _lnl(H, eb, lc, exptime, g, a, P, host, twin, window):
    t, f, s = lc
    o = np.argsort(t, kind="stable")
    t, f = np.ascontiguousarray(t[o]), np.ascontiguousarray(f[o])
    body = g["R_EB"] if eb else g["R_p"]
    cols = [np.ascontiguousarray(x, dtype=np.float64) for x in (
        body, g["EB_fluxratio"], P, g["inc"], a, g["R_s"], g["u1"], g["u2"], g["ecc"], g["argp"],
        g["cfr"])]
    out = np.empty(body.size)
    st = np.zeros(3, dtype=np.int64)
    H.hc_lnl(eb, t.size, _p(t), _p(f), s, exptime, 20, body.size, *[_p(c) for c in cols],
             int(host), int(twin), int(window), _p(out), st.ctypes.data_as(I64))
    return out, st


def test_occultation_and_separation_match_oracle(hc, golden):
    g = golden("model.npz")
    got = np.array([hc.hc_occult_quad(z, k, 0.4, 0.25) for z, k in zip(g["z"], g["k"])])
    # at the contact points (|z + k - 1| or |z - k| tiny) the acos arguments sit next to +-1
    # and the oracle's own k*k - z*z cancels: rounding is amplified to ~1e-13 there
    contact = (np.abs(g["z"] + g["k"] - 1) < 1e-5) | (np.abs(g["z"] - g["k"]) < 1e-3)
    np.testing.assert_allclose(got[~contact], g["flux"][~contact], rtol=0, atol=3e-14)
    np.testing.assert_allclose(got[contact], g["flux"][contact], rtol=0, atol=5e-13)
    zz = np.array([hc.hc_z(*x) for x in zip(g["t"], g["p"], g["a"], g["inc"], g["e"], g["w"])])
    # the cancellation in 1 - sin^2(w+f) sin^2 i amplifies rounding by (a/R*)^2 / z
    tol = 1e-15 * g["a"] ** 2 / np.maximum(np.abs(g["zsep"]), 1e-3) + 1e-14
    assert np.all(np.abs(zz - g["zsep"]) <= tol * np.maximum(1.0, np.abs(g["zsep"])))


@pytest.mark.parametrize("tag,lcname,star", [("toi465", "TOI465_01_lightcurve.csv", TOI465),
                                             ("kepler10b", "Kepler10b_lightcurve.csv", KEP10)])
def test_windowed_likelihood_matches_reference_fixtures(hc, golden, tag, lcname, star):
    lc = load_lc(lcname)
    g = golden("l1_%s.npz" % tag)
    exptime = float(g["exptime"])
    for host in (0, 1):
        for window in (1, 0):
            got, st = _lnl(hc, 0, lc, exptime, g, g["a"], g["P_orb"], host, 0, window)
            np.testing.assert_allclose(got, g["tp/%d" % host], rtol=1e-11)
            if window:
                assert st[1] > 0 and st[0] < got.size * lc[0].size   # windows really skip stamps
        got, _ = _lnl(hc, 1, lc, exptime, g, g["a"] * 1.2, g["P_orb"], host, 0, 1)
        want = g["eb/%d" % host]
        assert np.array_equal(np.isinf(got), np.isinf(want))
        fin = np.isfinite(want)
        np.testing.assert_allclose(got[fin], want[fin], rtol=1e-11)
        got, _ = _lnl(hc, 1, lc, exptime, g, g["a"] * 1.2 * 2 ** (2 / 3), 2 * g["P_orb"], host,
                      1, 1)
        np.testing.assert_allclose(got, g["twin/%d" % host], rtol=1e-11)


def test_window_falls_back_for_wide_or_wrapped_orbits(hc):
    """Short periods (window images reach the light curve), e >= 0.95 (table clamp) and
    grazing-orbit arcs must take the evaluate-everything path and still agree."""
    rng = np.random.default_rng(9)
    n = 64
    t = np.sort(rng.uniform(-1.5, 1.5, 400))
    f = 1 + rng.normal(0, 1e-3, t.size)
    g = dict(R_p=rng.uniform(5, 20, n), R_EB=rng.uniform(0.1, 1, n),
             EB_fluxratio=rng.uniform(0.01, 0.4, n), inc=rng.uniform(80, 90, n),
             R_s=np.full(n, 1.0), u1=np.full(n, 0.4), u2=np.full(n, 0.2),
             ecc=np.concatenate([rng.uniform(0, 0.6, n // 2), rng.uniform(0.93, 0.97, n // 2)]),
             argp=rng.uniform(0, 360, n), cfr=np.full(n, 0.1))
    P = rng.uniform(0.4, 2.5, n)
    a = 6.957e10 * rng.uniform(2.5, 8, n)
    want = coracle.lnL_TP_p(t, f, 1e-3, g["R_p"], P, g["inc"], a, g["R_s"], g["u1"], g["u2"],
                            g["ecc"], g["argp"], g["cfr"], False, 0.00139, 20)
    got, st = _lnl(hc, 0, (t, f, 1e-3), 0.00139, g, a, P, 0, 0, 1)
    np.testing.assert_allclose(got, want, rtol=1e-10)
