"""Randomised parity over the whole input domain: light curves that span several orbital
periods (window images reach the data), single-stamp and ragged curves, eccentricities up to
0.99 (beyond the orbit table), radius ratios from 1e-3 to 3, a/R* from 1.6 to 200, exposures up
to 0.2 d with 1/3/20 sub-exposures, both dilution conventions.  Tolerance: 1e-9 relative on
every draw's 0.5*chi^2, +inf pattern of the secondary-depth cut exact.

CPU: the device model's structure (tests/hostcheck) against the C oracle.
GPU: the kernels through the C ABI against the C oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import coracle

RSUN, REARTH = 6.957e10, 6.3781e8
D = ctypes.POINTER(ctypes.c_double)
I64 = ctypes.POINTER(ctypes.c_int64)


def _trial(rng, n):
    npts = int(rng.integers(1, 400))
    span = 10 ** rng.uniform(-1.5, 1.3)
    t = np.sort(rng.uniform(-span, span, npts))
    f = 1 + rng.normal(0, 1e-3, npts)
    exptime = float(rng.choice([0.0, 0.00139, 0.02, 0.2]))
    ns = int(rng.choice([1, 3, 20]))
    ars = 10 ** rng.uniform(0.2, 2.3, n)
    Rs = rng.uniform(0.3, 2, n)
    k = 10 ** rng.uniform(-3, 0.5, n)
    d = dict(P=10 ** rng.uniform(-0.5, 2, n), a=ars * Rs * RSUN, Rs=Rs,
             ecc=np.where(rng.random(n) < 0.3, 0.0, rng.uniform(0, 0.99, n)),
             argp=rng.uniform(0, 360, n),
             inc=np.degrees(np.arccos(np.minimum(rng.random(n) * 2.5 / ars, 1.0))),
             Rp=k * Rs * RSUN / REARTH, REB=k * Rs, u1=rng.uniform(0, 0.7, n),
             u2=rng.uniform(0, 0.3, n), cfr=rng.uniform(0.001, 0.9, n),
             fr=rng.uniform(1e-4, 0.6, n))
    return t, f, 1e-3, exptime, ns, d


def _compare(got, want, what):
    assert np.array_equal(np.isinf(got), np.isinf(want)), what
    assert np.array_equal(np.isnan(got), np.isnan(want)), what
    ok = np.isfinite(want)
    if ok.any():
        np.testing.assert_allclose(got[ok], want[ok], rtol=1e-9, atol=0, err_msg=what)


def _run(evaluate, seed, trials, n):
    rng = np.random.default_rng(seed)
    for trial in range(trials):
        t, f, s, ex, ns, d = _trial(rng, n)
        tp = (d["Rp"], d["P"], d["inc"], d["a"], d["Rs"], d["u1"], d["u2"], d["ecc"], d["argp"],
              d["cfr"])
        eb = (d["REB"], d["fr"]) + tp[1:]
        for host in (False, True):
            _compare(evaluate("tp", t, f, s, ex, ns, tp, host, False),
                     coracle.lnL_TP_p(t, f, s, *tp, host, ex, ns), ("tp", seed, trial, host))
            for twin in (False, True):
                fn = coracle.lnL_EB_twin_p if twin else coracle.lnL_EB_p
                _compare(evaluate("eb", t, f, s, ex, ns, eb, host, twin),
                         fn(t, f, s, *eb, host, ex, ns), ("eb", seed, trial, host, twin))


@pytest.fixture(scope="module")
def hostcheck():
    d = os.path.join(ROOT, "tests", "hostcheck")
    subprocess.check_call(["make", "-C", d, "-s"])
    H = ctypes.CDLL(os.path.join(d, "libhostcheck.so"))
    H.hc_lnl.argtypes = ([ctypes.c_int, ctypes.c_int64, D, D, ctypes.c_double, ctypes.c_double,
                          ctypes.c_int, ctypes.c_int64] + [D] * 11 + [ctypes.c_int] * 3 + [D, I64])
    return H


@pytest.mark.parametrize("seed", [1, 2])
def test_device_model_structure_on_cpu(hostcheck, seed):
    def evaluate(kind, t, f, s, ex, ns, cols, host, twin):
        p = lambda a: a.ctypes.data_as(D)  # noqa: E731
        if kind == "tp":
            cols = (cols[0], cols[0]) + cols[1:]            # unused EB flux-ratio slot
        arrs = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
        out = np.empty(arrs[0].size)
        st = np.zeros(3, dtype=np.int64)
        hostcheck.hc_lnl(int(kind == "eb"), t.size, p(t), p(f), s, ex, ns, out.size,
                         *[p(a) for a in arrs], int(host), int(twin), 1, p(out),
                         st.ctypes.data_as(I64))
        return out
    _run(evaluate, seed, trials=25, n=30)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_kernels_through_the_c_abi(gpu_engine, seed):
    from triceratops_b200 import likelihoods as lk

    def evaluate(kind, t, f, s, ex, ns, cols, host, twin):
        if kind == "tp":
            return lk.lnL_TP_p(t, f, s, *cols, host, ex, ns)
        fn = lk.lnL_EB_twin_p if twin else lk.lnL_EB_p
        return fn(t, f, s, *cols, host, ex, ns)
    _run(evaluate, seed, trials=30, n=200)
