"""The oracle itself: known-answer checks of the restated transit model's mathematics, the C
restatement against the numba one, and both against the committed fixtures (which were produced
by the reference's own likelihoods.py running over the restated model).

PARITY UNPINNED with respect to real pytransit==2.2 (absent): these tests
pin the mathematics (uniform-source closed form, brute-force quadrature, scipy elliptic
integrals) and the oracle's self-consistency, not PyTransit's rounding."""
import numpy as np
import pytest
from scipy import special

from oracle import coracle, quadmodel as qm


def test_hastings_polynomials_match_scipy_to_2e8():
    ks = np.linspace(0.01, 0.999, 400)
    assert max(abs(qm.ellk(k) - special.ellipk(k * k)) for k in ks) < 2e-8
    assert max(abs(qm.ellec(k) - special.ellipe(k * k)) for k in ks) < 2e-8


def test_bulirsch_third_kind_matches_carlson():
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = 10 ** rng.uniform(-2, 4)
        k = rng.uniform(0.01, 0.99)
        # Pi(-n | k^2) = RF(0, 1-k^2, 1) - n/3 ... with the (1 + n sin^2) sign convention
        rf = special.elliprf(0.0, 1 - k * k, 1.0)
        rj = special.elliprj(0.0, 1 - k * k, 1.0, 1.0 + n)
        want = rf - n / 3.0 * rj
        assert abs(qm.ellpicb(n, k) - want) < 5e-13 * max(1.0, abs(want))


def _uniform_source(z, k):
    """1 - (overlap area of two circles)/pi : closed form for u1 = u2 = 0."""
    if z >= 1 + k:
        return 1.0
    if z <= abs(1 - k):
        return 1.0 - min(k * k, 1.0)
    k0 = np.arccos((k * k + z * z - 1) / (2 * k * z))
    k1 = np.arccos((1 - k * k + z * z) / (2 * z))
    area = k * k * k0 + k1 - 0.5 * np.sqrt(max(4 * z * z - (1 + z * z - k * k) ** 2, 0.0))
    return 1.0 - area / np.pi


def test_uniform_source_limit_is_circle_overlap():
    rng = np.random.default_rng(1)
    for _ in range(2000):
        k = rng.uniform(0.01, 1.6)
        z = rng.uniform(0, 2.8)
        if abs(z - k) < 1e-3:
            continue
        assert abs(qm.eval_quad(z, k, 0.0, 0.0) - _uniform_source(z, k)) < 1e-7


def _brute(z, k, u1, u2, n=1500):
    r = (np.arange(n) + 0.5) / n
    th = (np.arange(4 * n) + 0.5) / (4 * n) * 2 * np.pi
    R, T = np.meshgrid(r, th, indexing="ij")
    mu = np.sqrt(1 - R ** 2)
    inten = 1 - u1 * (1 - mu) - u2 * (1 - mu) ** 2
    wgt = R
    occ = ((R * np.cos(T) - z) ** 2 + (R * np.sin(T)) ** 2) < k * k
    return 1 - np.sum(inten * wgt * occ) / np.sum(inten * wgt)


@pytest.mark.parametrize("z,k", [(0.0, 0.1), (0.3, 0.1), (0.85, 0.1), (0.95, 0.1), (1.05, 0.1),
                                 (0.3, 0.5), (0.5, 0.5), (0.7, 0.6), (1.2, 0.9), (0.9, 1.5),
                                 (2.0, 1.5), (0.6, 1.2), (0.2, 1.5)])
def test_limb_darkened_flux_matches_quadrature(z, k):
    assert abs(qm.eval_quad(z, k, 0.4, 0.25) - _brute(z, k, 0.4, 0.25)) < 1e-5


def test_total_and_no_occultation():
    assert qm.eval_quad(0.2, 1.5, 0.4, 0.25) == 0.0
    assert qm.eval_quad(1.2, 0.1, 0.4, 0.25) == 1.0
    assert qm.eval_quad(-0.3, 0.1, 0.4, 0.25) == 1.0     # far side of the orbit


def test_circular_orbit_is_time_symmetric_and_centred():
    es, ms, tae = qm.orbit_table()
    a, inc, p = 11.3, np.radians(88.0), 3.8
    z0 = qm.z_ip(0.0, 0.0, p, a, inc, 0.0, 0.0, es, ms, tae)
    assert abs(z0 - a * np.cos(inc)) < 1e-12              # impact parameter at mid-transit
    for t in (0.01, 0.05, 0.3):
        zp = qm.z_ip(t, 0.0, p, a, inc, 0.0, 0.0, es, ms, tae)
        zm = qm.z_ip(-t, 0.0, p, a, inc, 0.0, 0.0, es, ms, tae)
        assert abs(zp - zm) < 1e-9


def test_table_true_anomaly_close_to_kepler_solution():
    es, ms, tae = qm.orbit_table()
    rng = np.random.default_rng(2)
    for _ in range(300):
        e, w, p = rng.uniform(0, 0.6), rng.uniform(0, 6.28), rng.uniform(1, 20)
        t = rng.uniform(-0.5, 0.5) * p
        ta = qm.ta_ip(t, 0.0, p, e, w, es, ms, tae)
        off = qm.mean_anomaly_offset(e, w)
        ma = (2 * np.pi * (t - (0.0 - off * p / (2 * np.pi))) / p) % (2 * np.pi)
        exact = qm.ta_newton(ma, e) % (2 * np.pi)
        d = (ta - exact + np.pi) % (2 * np.pi) - np.pi
        assert abs(d) < 2e-4


def test_c_restatement_equals_numba_restatement(golden):
    g = golden("model.npz")
    got = np.array([coracle.eval_quad(z, k, 0.4, 0.25) for z, k in zip(g["z"], g["k"])])
    np.testing.assert_allclose(got, g["flux"], rtol=0, atol=2e-15)
    zz = np.array([coracle.lib().tro_z(*x) for x in zip(g["t"], g["p"], g["a"], g["inc"], g["e"],
                                                          g["w"])])
    np.testing.assert_allclose(zz, g["zsep"], rtol=1e-13, atol=1e-13)
    _, _, tae = coracle.orbit_table()
    np.testing.assert_array_equal(tae[::17, ::31], g["tae_sample"])


@pytest.mark.parametrize("tag,lcname", [("toi465", "TOI465_01_lightcurve.csv"),
                                        ("kepler10b", "Kepler10b_lightcurve.csv")])
def test_c_oracle_reproduces_reference_likelihood_fixtures(golden, tag, lcname):
    """Fixtures = reference likelihoods.py:443-587 executed over the restated model."""
    from conftest import load_lc
    t, f, s = load_lc(lcname)
    g = golden("l1_%s.npz" % tag)
    exptime = float(g["exptime"])
    for host in (0, 1):
        got = coracle.lnL_TP_p(t, f, s, g["R_p"], g["P_orb"], g["inc"], g["a"], g["R_s"],
                               g["u1"], g["u2"], g["ecc"], g["argp"], g["cfr"], host, exptime, 20)
        np.testing.assert_allclose(got, g["tp/%d" % host], rtol=1e-12)
        got = coracle.lnL_EB_p(t, f, s, g["R_EB"], g["EB_fluxratio"], g["P_orb"], g["inc"],
                               g["a"] * 1.2, g["R_s"], g["u1"], g["u2"], g["ecc"], g["argp"],
                               g["cfr"], host, exptime, 20)
        want = g["eb/%d" % host]
        assert np.array_equal(np.isinf(got), np.isinf(want))
        fin = np.isfinite(want)
        np.testing.assert_allclose(got[fin], want[fin], rtol=1e-12)
        got = coracle.lnL_EB_twin_p(t, f, s, g["R_EB"], g["EB_fluxratio"], 2 * g["P_orb"],
                                    g["inc"], g["a"] * 1.2 * 2 ** (2 / 3), g["R_s"], g["u1"],
                                    g["u2"], g["ecc"], g["argp"], g["cfr"], host, exptime, 20)
        np.testing.assert_allclose(got, g["twin/%d" % host], rtol=1e-12)
