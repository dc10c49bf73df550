"""A whole calc_probs (all 15 target rows, contrast curve) through the CPU stand-in engine with
the C preparation blocks and with the numpy statements: the same evidences, probabilities and
best-draw tables, bit for bit."""
import os

import numpy as np

from triceratops_b200 import _blocks

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _run(oracle_engine, t, f, s, N):
    from conftest import TOI465
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    stars = synth.stars_table(9, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"], TOI465["M"],
                              TOI465["R"], TOI465["Teff"], TOI465["plx"], n_neighbours=0)
    tgt = target(9, stars=stars, trilegal_fname=os.path.join(GOLD, "trilegal_synth.csv"))
    np.random.seed(4)
    tgt.calc_probs(t, f, s, TOI465["P"], N=N, parallel=True, verbose=0, filt="K",
                   contrast_curve_file=os.path.join(GOLD, "TOI465_01_contrastcurve.csv"))
    return tgt


def test_calc_probs_is_identical_with_c_blocks_and_numpy_statements(oracle_engine, toi465_lc,
                                                                    monkeypatch):
    assert _blocks.available()
    t, f, s = toi465_lc
    t, f = t[::12], f[::12]                # (the CPU stand-in evaluates every surviving draw)
    N = _blocks.MIN_N + 1
    used = []
    real = _blocks.run

    def spy(kind, n, **kw):
        out = real(kind, n, **kw)
        used.append(out is not None)
        return out
    monkeypatch.setattr(_blocks, "run", spy)
    a = _run(oracle_engine, t, f, s, N)
    assert len(used) == 10 and all(used)
    monkeypatch.setattr(_blocks, "run", lambda *x, **k: None)
    b = _run(oracle_engine, t, f, s, N)
    assert np.array_equal(a.lnZ, b.lnZ, equal_nan=True)
    assert a.FPP == b.FPP and a.NFPP == b.NFPP
    for col in a.probs.columns:
        x, y = a.probs[col].values, b.probs[col].values
        if x.dtype.kind == "f":
            assert np.array_equal(x, y, equal_nan=True), col
        else:
            assert list(x) == list(y), col
