"""The scenarios' element-wise preparation runs chunk by chunk over the host threads
(_hostpar.pmap_block).  Chunking must not change a single bit of what is handed to the engine."""
import os

import numpy as np
import pytest

import triceratops_b200.marginal_likelihoods as ml
from triceratops_b200 import _hostpar

HERE = os.path.dirname(os.path.abspath(__file__))
N = 300_000


def _record(monkeypatch):
    calls = []
    monkeypatch.setattr(ml, "_run_tp", lambda *a, **k: calls.append(("tp", a, k)))
    monkeypatch.setattr(ml, "_run_eb", lambda *a, **k: calls.append(("eb", a, k)) or (None, None))
    return calls


def _same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a), np.asarray(b)
        return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
    return a == b or (a is None and b is None)


def _scenarios(star):
    lc = star["lc"]
    common = dict(N=N, parallel=True, exptime=lc[3], nsamples=lc[4])
    tfs = (lc[0], lc[1], lc[2])
    S = star
    yield "TTP", lambda: ml.lnZ_TTP(*tfs, S["P"], S["M"], S["R"], S["Teff"], S["Z"], **common)
    yield "TEB", lambda: ml.lnZ_TEB(*tfs, S["P"], S["M"], S["R"], S["Teff"], S["Z"], **common)
    for name in ("PTP", "PEB", "STP", "SEB"):
        for cc, filt in ((None, "TESS"), (S["cc"], "J")):
            yield name + ("cc" if cc else ""), (lambda name=name, cc=cc, filt=filt: getattr(
                ml, "lnZ_" + name)(*tfs, S["P"], S["M"], S["R"], S["Teff"], S["Z"], S["plx"],
                                   cc, filt, **common))
    for name in ("DTP", "DEB"):
        for cc, filt in ((None, "TESS"), (S["cc"], "K")):
            yield name + ("cc" if cc else ""), (lambda name=name, cc=cc, filt=filt: getattr(
                ml, "lnZ_" + name)(*tfs, S["P"], S["M"], S["R"], S["Teff"], S["Z"], S["Tmag"],
                                   S["Jmag"], S["Hmag"], S["Kmag"], S["trilegal"], cc, filt,
                                   **common))
    for name in ("BTP", "BEB"):
        for cc, filt in ((None, "TESS"), (S["cc"], "H")):
            yield name + ("cc" if cc else ""), (lambda name=name, cc=cc, filt=filt: getattr(
                ml, "lnZ_" + name)(*tfs, S["P"], S["M"], S["R"], S["Teff"], S["Tmag"], S["Jmag"],
                                   S["Hmag"], S["Kmag"], S["trilegal"], cc, filt, **common))


@pytest.fixture
def star():
    t = np.linspace(-0.2, 0.2, 40)
    return {"lc": (t, np.ones_like(t), 1e-3, 0.00139, 3), "P": 4.2, "M": 0.93, "R": 0.95,
            "Teff": 5400.0, "Z": 0.05, "plx": 8.1, "Tmag": 10.3, "Jmag": 9.6, "Hmag": 9.2,
            "Kmag": 9.1, "trilegal": os.path.join(HERE, "golden", "trilegal_synth.csv"),
            "cc": os.path.join(HERE, "golden", "TOI465_01_contrastcurve.csv")}


def test_pmap_block_pieces():
    if _hostpar._pool is None:
        pytest.skip("single host thread")
    n = 400_001
    x = np.random.default_rng(3).random(n)
    idx = np.arange(n)

    def fn(a, k, scale):
        return a * scale, None, a > 0.5, k

    y, none, m, k = _hostpar.pmap_block(fn, n, x, idx, 3.0)
    assert none is None and np.array_equal(y, x * 3.0) and np.array_equal(m, x > 0.5)
    assert np.array_equal(k, idx) and k.dtype == idx.dtype and m.dtype == bool
    assert len(_hostpar.block_chunks(n)) > 1


def test_block_failure_propagates():
    if _hostpar._pool is None:
        pytest.skip("single host thread")

    def fn(a):
        raise ValueError("bad chunk")
    with pytest.raises(ValueError, match="bad chunk"):
        _hostpar.pmap_block(fn, 400_000, np.zeros(400_000))
    # the worker threads are usable again (their inline flag was reset)
    assert len(_hostpar.pmap(lambda a: a + 1, 400_000, np.zeros(400_000))) > 1


def test_chunked_preparation_is_bit_identical(star, monkeypatch):
    if _hostpar._pool is None:
        pytest.skip("single host thread")
    for f in (star["trilegal"], star["cc"]):
        if not os.path.exists(f):
            pytest.skip("fixture %s missing" % f)
    calls = _record(monkeypatch)
    monkeypatch.setattr(ml._dispatch, "use_lightcurve", lambda *a, **k: None)
    monkeypatch.setattr(ml._blocks, "run", lambda *a, **k: None)     # the numpy path is under test
    for name, run in _scenarios(star):
        np.random.seed(11)
        del calls[:]
        run()
        chunked = list(calls)
        pool = _hostpar._pool
        monkeypatch.setattr(_hostpar, "_pool", None)
        try:
            np.random.seed(11)
            del calls[:]
            run()
            whole = list(calls)
        finally:
            monkeypatch.setattr(_hostpar, "_pool", pool)
        assert len(chunked) == len(whole) >= 1, name
        for (k1, a1, kw1), (k2, a2, kw2) in zip(chunked, whole):
            assert k1 == k2 and len(a1) == len(a2) and kw1 == kw2, name
            for i, (u, v) in enumerate(zip(a1, a2)):
                assert _same(u, v), "%s: argument %d differs between chunked and whole" % (name, i)
