"""Per-point errors (an extension of the reference's scalar sigma):
tri_set_lightcurve_err through the engine and through calc_probs.

chi^2 = sum_j (flux_j - model_j)^2 / sigma_j^2; the Gaussian constant (applied once per light
curve, marginal_likelihoods.py:130) and the secondary-depth cut (likelihoods.py:535) use
sigma = mean(sigma_j).  Checked against the oracle's model light curves weighted in numpy, and
against the scalar call for constant errors."""
import numpy as np
import pytest

from conftest import TOI465, draw_eb_columns, draw_tp_columns

pytestmark = pytest.mark.gpu


def test_constant_errors_equal_the_scalar_call(gpu_engine, toi465_lc):
    t, f, s = toi465_lc
    N = 20000
    cols = draw_tp_columns(N, 3)
    gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    a = gpu_engine.eval_tp(N, **cols, want_mask=True)
    gpu_engine.set_lightcurve(t, f, np.full(t.size, s), 0.00139, 20)
    b = gpu_engine.eval_tp(N, **cols, want_mask=True)
    gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    assert np.array_equal(a.mask, b.mask)
    fin = np.isfinite(a.lnL)
    np.testing.assert_allclose(b.lnL[fin], a.lnL[fin], rtol=1e-13)
    assert abs(a.lnZ - b.lnZ) < 1e-9
    ecols = draw_eb_columns(N, 4)
    a = gpu_engine.eval_eb(N, **ecols)
    gpu_engine.set_lightcurve(t, f, np.full(t.size, s), 0.00139, 20)
    b = gpu_engine.eval_eb(N, **ecols)
    gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    for x, y in zip(a, b):
        assert np.array_equal(np.isfinite(x.lnL), np.isfinite(y.lnL))
        fin = np.isfinite(x.lnL)
        np.testing.assert_allclose(y.lnL[fin], x.lnL[fin], rtol=1e-13)


def test_weighted_chi2_against_oracle_models(gpu_engine, toi465_lc):
    from oracle import coracle
    from oracle.engine_port import Rearth, Rsun, tp_mask
    t, f, s = toi465_lc
    rng = np.random.default_rng(8)
    err = s * rng.uniform(0.5, 2.0, t.size)
    N = 4000
    cols = draw_tp_columns(N, 5)
    gpu_engine.set_lightcurve(t, f, err, 0.00139, 20)
    try:
        g = gpu_engine.eval_tp(N, **cols, want_mask=True)
    finally:
        gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    mask, a = tp_mask(N, cols["rp"], cols["P_orb"], cols["inc"], cols["ecc"], cols["argp"],
                      cols["mtot"], cols["rhost"])
    assert np.array_equal(g.mask, mask)
    idx = np.flatnonzero(mask)[:300]
    sigma = err.mean()
    const = -0.5 * np.log(2 * np.pi) - np.log(sigma)
    want = np.empty(idx.size)
    for n, i in enumerate(idx):
        model = coracle.model(t, cols["rp"][i] * Rearth / (cols["rhost"] * Rsun), cols["P_orb"],
                              a[i] / (cols["rhost"] * Rsun), np.radians(cols["inc"][i]),
                              cols["ecc"][i], np.radians(90 - cols["argp"][i]), cols["u1"],
                              cols["u2"], 0.00139, 20)
        want[n] = const - 0.5 * np.sum((f - model) ** 2 / err ** 2)
    np.testing.assert_allclose(g.lnL[idx], want, rtol=1e-9)


def test_calc_probs_accepts_an_error_array(gpu_engine, toi465_lc, trilegal_file):
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    t, f, s = toi465_lc
    stars = synth.stars_table(9, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"], TOI465["M"],
                              TOI465["R"], TOI465["Teff"], TOI465["plx"], n_neighbours=0)
    out = []
    for err in (s, np.full(t.size, s)):
        tgt = target(9, stars=stars, trilegal_fname=trilegal_file)
        np.random.seed(1)
        tgt.calc_probs(t, f, err, TOI465["P"], N=5000, parallel=True, verbose=0)
        out.append(tgt)
    np.testing.assert_allclose(out[1].lnZ, out[0].lnZ, rtol=0, atol=1e-9)
    np.testing.assert_allclose(out[1].probs.prob.values, out[0].probs.prob.values, atol=1e-9)
