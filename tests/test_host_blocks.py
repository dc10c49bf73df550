"""csrc/host_blocks.c runs a scenario's element-wise preparation in one GIL-free call.  It must
hand the engine exactly the bits the numpy statements of marginal_likelihoods.py produce: every
scenario, with and without contrast curve, every band, flat and default priors, MOLUSC table,
every primary-mass regime, both eccentricity laws, ragged chunk tails."""
import os

import numpy as np
import pytest

import triceratops_b200.marginal_likelihoods as ml
from triceratops_b200 import _blocks

HERE = os.path.dirname(os.path.abspath(__file__))
TRI = os.path.join(HERE, "golden", "trilegal_synth.csv")
CC = os.path.join(HERE, "golden", "TOI465_01_contrastcurve.csv")
N = 40_961          # 5 chunks of 8192 + 1


@pytest.fixture(autouse=True)
def _c_path_is_live():
    assert _blocks.available(), "host_blocks.c unavailable or its self-check failed"


def _record(monkeypatch):
    calls = []
    monkeypatch.setattr(ml, "_run_tp", lambda *a, **k: calls.append(("tp", a, k)))
    monkeypatch.setattr(ml, "_run_eb", lambda *a, **k: calls.append(("eb", a, k)) or (None, None))
    monkeypatch.setattr(ml._dispatch, "use_lightcurve", lambda *a, **k: None)
    return calls


def _same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a), np.asarray(b)
        return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
    return a == b or (a is None and b is None)


def _both_ways(monkeypatch, calls, run, label):
    used = []
    real_run = _blocks.run

    def spy(kind, n, **kw):
        out = real_run(kind, n, **kw)
        used.append((kind, out is not None))
        return out
    with monkeypatch.context() as mp:
        mp.setattr(_blocks, "run", spy)
        np.random.seed(11)
        del calls[:]
        run()
        c_side = list(calls)
    assert used and all(ok for _, ok in used), "%s: the C path was not taken: %s" % (label, used)
    with monkeypatch.context() as mp:
        # the numpy side: one whole-array evaluation, as the reference does it
        mp.setattr(_blocks, "run", lambda *a, **k: None)
        mp.setattr(ml._hostpar, "pmap_block", lambda fn, n, *arrays: fn(*arrays))
        np.random.seed(11)
        del calls[:]
        run()
        np_side = list(calls)
    assert len(c_side) == len(np_side) >= 1, label
    for (k1, a1, kw1), (k2, a2, kw2) in zip(c_side, np_side):
        assert k1 == k2 and len(a1) == len(a2) and kw1 == kw2, label
        for i, (u, v) in enumerate(zip(a1, a2)):
            assert _same(u, v), "%s: argument %d of _run_%s differs (C vs numpy)" % (label, i, k1)


T = np.linspace(-0.2, 0.2, 40)
LC = (T, np.ones_like(T), 1e-3)
STARS = [  # (M_s, R_s, Teff, Z, P_orb)
    (0.93, 0.95, 5400.0, 0.05, 4.2),
    (1.31, 1.6, 6300.0, -0.2, 17.0),
    (0.40, 0.41, 3600.0, 0.0, 2.1),
    (0.24, 0.27, 3300.0, 0.1, 11.0),
    (0.09, 0.12, 2900.0, 0.0, 0.9),
]


@pytest.mark.parametrize("star", STARS)
@pytest.mark.parametrize("flat", [False, True])
def test_target_scenarios(monkeypatch, star, flat):
    M, R, Te, Z, P = star
    calls = _record(monkeypatch)
    kw = dict(N=N, parallel=True, flatpriors=flat)
    _both_ways(monkeypatch, calls, lambda: ml.lnZ_TTP(*LC, P, M, R, Te, Z, **kw), "TTP")
    _both_ways(monkeypatch, calls, lambda: ml.lnZ_TEB(*LC, P, M, R, Te, Z, **kw), "TEB")


@pytest.mark.parametrize("star", STARS[:4])
@pytest.mark.parametrize("cc,filt", [(None, "TESS"), (CC, "TESS"), (CC, "Vis"), (CC, "J"),
                                     (CC, "H"), (CC, "K")])
@pytest.mark.parametrize("name", ["PTP", "PEB", "STP", "SEB"])
def test_bound_companion_scenarios(monkeypatch, star, cc, filt, name):
    M, R, Te, Z, P = star
    calls = _record(monkeypatch)
    fn = getattr(ml, "lnZ_" + name)
    for plx in (8.1, float("nan")):
        _both_ways(monkeypatch, calls,
                   lambda: fn(*LC, P, M, R, Te, Z, plx, cc, filt, N=N, parallel=True),
                   "%s cc=%s filt=%s plx=%s" % (name, bool(cc), filt, plx))


@pytest.mark.parametrize("name", ["PTP", "PEB", "STP", "SEB"])
def test_bound_companion_scenarios_with_molusc_table(monkeypatch, tmp_path, name):
    import pandas as pd
    rng = np.random.default_rng(5)
    n = 30_000
    p = tmp_path / "molusc.csv"
    pd.DataFrame({"semi-major axis(AU)": rng.uniform(1, 200, n),
                  "eccentricity": rng.uniform(0, 0.9, n),
                  "mass ratio": rng.uniform(0.01, 1.0, n)}).to_csv(p, index=False)
    calls = _record(monkeypatch)
    fn = getattr(ml, "lnZ_" + name)
    M, R, Te, Z, P = STARS[0]
    _both_ways(monkeypatch, calls,
               lambda: fn(*LC, P, M, R, Te, Z, 8.1, CC, "J", N=N, parallel=True,
                          molusc_file=str(p)), name + " molusc")


@pytest.mark.parametrize("star", [STARS[0], STARS[1], STARS[3]])
@pytest.mark.parametrize("cc,filt", [(None, "TESS"), (CC, "TESS"), (CC, "J"), (CC, "H"),
                                     (CC, "K")])
@pytest.mark.parametrize("name", ["DTP", "DEB", "BTP", "BEB"])
def test_background_scenarios(monkeypatch, star, cc, filt, name):
    M, R, Te, Z, P = star
    calls = _record(monkeypatch)
    fn = getattr(ml, "lnZ_" + name)
    mags = (10.3, 9.6, 9.2, 9.1)
    for flat in ((False, True) if name.endswith("TP") else (False,)):
        if name[0] == "D":
            run = lambda: fn(*LC, P, M, R, Te, Z, *mags, TRI, cc, filt, N=N, parallel=True,  # noqa: E731
                             flatpriors=flat)
        else:
            run = lambda: fn(*LC, P, M, R, Te, *mags, TRI, cc, filt, N=N, parallel=True,  # noqa: E731
                             flatpriors=flat)
        _both_ways(monkeypatch, calls, run, "%s cc=%s filt=%s flat=%s" % (name, bool(cc), filt, flat))


def test_interp_port_matches_numpy_on_awkward_tables(monkeypatch):
    """numpy.interp as ported (guess-based search, exact-node and NaN rules), through the
    background prior: non-monotonic contrast tables, repeated nodes, a single node, queries on
    nodes, outside the table and NaN."""
    calls = _record(monkeypatch)
    rng = np.random.default_rng(9)
    import tempfile
    for k, (seps, cons) in enumerate([
            (np.linspace(0.05, 4, 60), np.sort(rng.uniform(0, 9, 60))),
            (np.linspace(0.05, 4, 60), rng.uniform(0, 9, 60)),                 # not monotonic
            (np.array([0.1, 0.5, 0.5, 2.0, 3.0]), np.array([1.0, 3.0, 3.0, 6.0, 7.5])),
            (np.array([1.5, 2.5]), np.array([4.0, 4.0])),
            (np.linspace(0.1, 3, 3), np.array([1.0, 2.0, 8.0]))]):
        with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as fh:
            np.savetxt(fh, np.column_stack([seps, cons]), delimiter=",")
        try:
            for name in ("DEB", "BEB", "PEB"):
                fn = getattr(ml, "lnZ_" + name)
                M, R, Te, Z, P = STARS[0]
                if name == "DEB":
                    run = lambda: fn(*LC, P, M, R, Te, Z, 10.3, 9.6, 9.2, 9.1, TRI, fh.name, "K",  # noqa: E731
                                     N=N, parallel=True)
                elif name == "BEB":
                    run = lambda: fn(*LC, P, M, R, Te, 10.3, 9.6, 9.2, 9.1, TRI, fh.name, "J",  # noqa: E731
                                     N=N, parallel=True)
                else:
                    run = lambda: fn(*LC, P, M, R, Te, Z, 8.1, fh.name, "H", N=N, parallel=True)  # noqa: E731
                _both_ways(monkeypatch, calls, run, "%s table %d" % (name, k))
        finally:
            os.unlink(fh.name)


def test_small_calls_and_disabled_path_use_numpy(monkeypatch):
    assert _blocks.run("TTP", 100, M_s=1.0, R_s=1.0, Teff=5000.0, x_inc=np.zeros(100),
                       x_w=np.zeros(100), x_rp=np.zeros(100)) is None
    monkeypatch.setattr(_blocks, "_state", False)
    assert _blocks.run("TTP", N, M_s=1.0, R_s=1.0, Teff=5000.0, x_inc=np.zeros(N),
                       x_w=np.zeros(N), x_rp=np.zeros(N)) is None


def test_random_stars_small_draw_counts(monkeypatch):
    """Many random targets (including the regime edges M = 0.1, 0.3, 0.45, 0.63, 1.0 Msun and
    periods on both sides of the P = 10 d eccentricity switch) at a draw count that leaves
    ragged chunks: every scenario, bit for bit."""
    monkeypatch.setattr(_blocks, "MIN_N", 1000)
    calls = _record(monkeypatch)
    rng = np.random.default_rng(2026)
    n = 9_001
    masses = [0.1, 0.3, 0.45, 0.63, 1.0, 0.2999999, 1.0000001] + list(rng.uniform(0.08, 2.5, 25))
    mags = (10.3, 9.6, 9.2, 9.1)
    for k, M in enumerate(masses):
        R = float(M ** 0.9 * rng.uniform(0.8, 1.2))
        Te = float(np.clip(5777 * M ** 0.55, 2900, 9500))
        Z = float(rng.uniform(-0.5, 0.4))
        P = float(rng.choice([0.7, 3.3, 9.999, 10.0, 10.001, 25.0]))
        plx = float(rng.choice([0.5, 8.1, 120.0]))
        cc, filt = [(None, "TESS"), (CC, "TESS"), (CC, "J"), (CC, "H"), (CC, "K"), (CC, "Vis")][k % 6]
        flat = bool(k % 3 == 0)
        kw = dict(N=n, parallel=True)
        tag = "M=%.4f P=%.3f %s" % (M, P, filt)
        _both_ways(monkeypatch, calls,
                   lambda: ml.lnZ_TTP(*LC, P, M, R, Te, Z, flatpriors=flat, **kw), "TTP " + tag)
        _both_ways(monkeypatch, calls, lambda: ml.lnZ_TEB(*LC, P, M, R, Te, Z, **kw), "TEB " + tag)
        for name in ("PTP", "PEB", "STP", "SEB"):
            fn = getattr(ml, "lnZ_" + name)
            _both_ways(monkeypatch, calls,
                       lambda: fn(*LC, P, M, R, Te, Z, plx, cc, filt, flatpriors=flat, **kw),
                       name + " " + tag)
        for name in ("DTP", "DEB"):
            fn = getattr(ml, "lnZ_" + name)
            _both_ways(monkeypatch, calls,
                       lambda: fn(*LC, P, M, R, Te, Z, *mags, TRI, cc, filt, flatpriors=flat, **kw),
                       name + " " + tag)
        for name in ("BTP", "BEB"):
            fn = getattr(ml, "lnZ_" + name)
            _both_ways(monkeypatch, calls,
                       lambda: fn(*LC, P, M, R, Te, *mags, TRI, cc, filt, flatpriors=flat, **kw),
                       name + " " + tag)


def test_period_range_inputs(monkeypatch):
    """P_orb given as a range: the periods are drawn too and the eccentricity law follows their
    mean (reference marginal_likelihoods.py:67-76)."""
    calls = _record(monkeypatch)
    M, R, Te, Z, _ = STARS[0]
    for P in (np.array([3.0, 12.0]), np.array([11.0, 30.0])):
        _both_ways(monkeypatch, calls, lambda: ml.lnZ_TEB(*LC, P, M, R, Te, Z, N=N, parallel=True),
                   "TEB range")
        _both_ways(monkeypatch, calls,
                   lambda: ml.lnZ_SEB(*LC, P, M, R, Te, Z, 8.1, CC, "J", N=N, parallel=True),
                   "SEB range")
