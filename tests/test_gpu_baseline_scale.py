"""Parity at BASELINE.json's own sizes (north_star gates: masks bit-exact, per-draw lnL to 1e-9
relative, lnZ and scenario probabilities to 1e-6), through the C ABI on the B200.

  config 1  N = 1e5, lnZ_TTP + lnZ_TEB: every surviving draw against the C oracle
  config 2  N = 1e6, the 12 engine calls of the 18-row calc_probs: masks of all N draws, lnL of
            a seeded subset of the survivors, device lnZ against the host log-mean-exp of the
            device lnL, device best-draw selection against a stable sort
  config 3  the same on Kepler-10b with 29.4-min exposures
  config 4  N = 1e7 host-drawn samples on the 20 000-stamp light curve, the same subset method
  calc_probs at N = 1e6: probabilities / FPP / NFPP against the same call routed to the oracle
            port (a full oracle pass: the slowest test of the suite, about two minutes of host
            CPU)
The input columns are what `target.calc_probs` itself produces for a numpy seed
(_workloads.record_calls), i.e. identical host-drawn sample arrays on both sides.
"""
import numpy as np
import pytest

import _parity
import _workloads

pytestmark = pytest.mark.gpu

SEED = 2026


def _check_config(gpu_engine, config, N, n_sub, lc=None):
    lc = lc if lc is not None else _workloads.lightcurve(config)
    calls, _ = _workloads.record_calls(config, N, SEED, lc)
    gates = []
    for k, call in enumerate(calls):
        g = _parity.check_call(gpu_engine, call, n_sub=n_sub, seed=k)
        _parity.assert_gates(g)
        gates.append(g)
    return _parity.merge(gates), calls


def test_config1_every_surviving_draw(gpu_engine):
    """BASELINE configs[0]: TOI-465.01, TP / EB / EBx2P only, N = 1e5 -- full oracle pass."""
    g, calls = _check_config(gpu_engine, 1, 100_000, None)
    assert len(calls) == 2 and g["n_checked"] == g["n_pass"] > 15_000


def test_config2_all_engine_calls_at_1e6(gpu_engine):
    """BASELINE configs[1]: the bench workload itself."""
    g, calls = _check_config(gpu_engine, 2, 1_000_000, 4000)
    assert len(calls) == 12 and g["n_checked"] > 50_000


def test_config3_kepler_long_cadence_at_1e6(gpu_engine):
    """BASELINE configs[2]: 30-min exposures (the centre probe skips far less)."""
    g, calls = _check_config(gpu_engine, 3, 1_000_000, 4000)
    assert len(calls) == 12 and g["n_checked"] > 40_000


def test_config4_host_sampler_at_1e7(gpu_engine):
    """BASELINE configs[3] in the mode that can be checked: numpy draws, N = 1e7, 20 000 stamps.
    Three of the twelve engine calls (TTP, TEB, BEB: scalar host, EB branches, per-draw host)
    keep the test's host memory and time bounded; bench.py --config 4 covers all twelve."""
    from triceratops_b200 import marginal_likelihoods as ml
    lc = _workloads.lightcurve(4)
    t, f, s = lc
    star = _workloads.SOLAR
    N = 10_000_000
    rec = _workloads.Recorder()
    tail = (N, True, "TESS", False, 0.00139, 20)
    base = (t, f, s, star["P"], star["M"], star["R"], star["Teff"])
    mags = (star["T"], star["J"], star["H"], star["K"])
    tri = _workloads.os.path.join(_workloads.GOLD, "trilegal_synth.csv")
    cc = _workloads.os.path.join(_workloads.GOLD, "TOI465_01_contrastcurve.csv")
    gates = []
    for k, fn in enumerate((lambda: ml.lnZ_TTP(*base, 0.0, *tail),
                            lambda: ml.lnZ_TEB(*base, 0.0, *tail),
                            lambda: ml.lnZ_BEB(*base, *mags, tri, cc, "K", *tail))):
        rec.calls.clear()
        with _workloads.patched_engine(rec):
            np.random.seed(SEED + k)
            fn()
        g = _parity.check_call(gpu_engine, rec.calls[0], n_sub=600, seed=k)
        _parity.assert_gates(g)
        gates.append(g)
    g = _parity.merge(gates)
    assert g["n_pass"] > 500_000 and g["n_checked"] >= 1800


def test_calc_probs_probabilities_at_1e6_against_oracle(gpu_engine):
    """The public call at the bench size on the CUDA engine and on the oracle port, same numpy
    seed: every scenario's lnZ, the probabilities, FPP and NFPP (north_star: 1e-6)."""
    import _oracle_engine
    lc = _workloads.lightcurve(2)
    N = 1_000_000
    gpu = _workloads.run_calc_probs(_workloads.make_target(2), 2, lc, N, SEED)
    ora = _oracle_engine.install()
    try:
        cpu = _workloads.run_calc_probs(_workloads.make_target(2), 2, lc, N, SEED)
    finally:
        _oracle_engine.uninstall()
    assert ora is not None and len(gpu.lnZ) == 18
    fin = np.isfinite(cpu.lnZ)
    assert np.array_equal(np.isfinite(gpu.lnZ), fin)
    np.testing.assert_allclose(gpu.lnZ[fin], cpu.lnZ[fin], rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(gpu.probs.prob.values, cpu.probs.prob.values, rtol=0, atol=1e-6)
    assert abs(gpu.FPP - cpu.FPP) < 1e-6 and abs(gpu.NFPP - cpu.NFPP) < 1e-6
    # the best draw of every scenario is the same draw
    for col in ("R_p", "inc", "ecc", "M_EB"):
        np.testing.assert_allclose(gpu.probs[col].values, cpu.probs[col].values, rtol=1e-9)
