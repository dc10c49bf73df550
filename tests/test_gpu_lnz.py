"""GPU parity at the scenario seam: the ten lnZ_* functions and calc_probs on the CUDA engine
against the fixtures produced by the reference's own marginal_likelihoods.py / triceratops.py
(lnZ and probabilities to 1e-6), and the fused kernels against the oracle stand-in on identical
draws (masks bit-exact, per-draw lnL to 1e-9)."""
import numpy as np
import pytest

from conftest import (KEP10, SCALAR_NAMES, TOI465, _Prefixed, check_against_golden, lnz_calls,
                      nearby_calls, scalar_calls, scalar_star)

import triceratops_b200.marginal_likelihoods as ml

pytestmark = pytest.mark.gpu

NAMES = ["TTP", "TEB", "PTP", "PTPcc", "PEB", "PEBcc", "STP", "STPcc", "SEB", "SEBcc", "DTP",
         "DTPcc", "DEB", "DEBcc", "BTP", "BTPcc", "BEB", "BEBcc"]


@pytest.mark.parametrize("name", NAMES)
def test_lnz_functions_match_reference_fixtures(name, gpu_engine, golden, toi465_lc,
                                                trilegal_file, contrast_file):
    g = golden("lnz_toi465.npz")
    calls = lnz_calls(TOI465, int(g["N"]), trilegal_file, contrast_file, toi465_lc)
    np.random.seed(int(g["seed"]))
    check_against_golden(name, calls[name](ml), g, lnz_atol=1e-6, arr_rtol=1e-9)


@pytest.mark.parametrize("name", NAMES)
def test_kepler_long_cadence(name, gpu_engine, golden, kepler10b_lc, trilegal_file, contrast_file):
    """BASELINE config 3: Kepler-10b, 30-min exposure supersampling, every scenario."""
    g = golden("lnz_kepler10b.npz")
    calls = lnz_calls(KEP10, int(g["N"]), trilegal_file, contrast_file, kepler10b_lc,
                      mission="Kepler", exptime=0.0204)
    np.random.seed(int(g["seed"]))
    check_against_golden(name, calls[name](ml), g, lnz_atol=1e-6, arr_rtol=1e-9)


@pytest.mark.parametrize("tag", ["toi465", "tight"])
@pytest.mark.parametrize("name", SCALAR_NAMES)
def test_parallel_false_matches_reference_scalar_loops(name, tag, gpu_engine, golden, toi465_lc,
                                                       trilegal_file, contrast_file):
    """parallel=False, the reference's default (triceratops.py:676): its per-draw loops over the
    scalar likelihoods, reproduced by the engine's scalar_loop flag."""
    g = golden("lnz_scalar.npz")
    star = scalar_star(g, tag)
    calls = scalar_calls(star, int(g["N"]), trilegal_file, contrast_file, toi465_lc)
    np.random.seed(int(g["seed"]))
    check_against_golden(name, calls[name](ml), _Prefixed(g, tag + "/"), lnz_atol=1e-6,
                         arr_rtol=1e-9)


def test_calc_probs_matches_reference_fixture(gpu_engine, golden, toi465_lc, trilegal_file,
                                              contrast_file):
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    g = golden("calc_probs.npz")
    t, f, s = toi465_lc
    stars = synth.stars_table(270380593, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"],
                              TOI465["M"], TOI465["R"], TOI465["Teff"], TOI465["plx"])
    tgt = target(270380593, stars=stars, trilegal_fname=trilegal_file)
    np.random.seed(int(g["seed"]))
    tgt.calc_probs(t, f, s, TOI465["P"], contrast_curve_file=contrast_file, filt="K",
                   N=int(g["N"]), parallel=True, verbose=0)
    assert list(tgt.probs.scenario.values) == list(g["scenario"])
    fin = np.isfinite(g["lnZ"])
    assert np.array_equal(np.isfinite(tgt.lnZ), fin)
    np.testing.assert_allclose(tgt.lnZ[fin], g["lnZ"][fin], rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(tgt.probs.prob.values, g["prob"], rtol=0, atol=1e-6)
    assert abs(tgt.FPP - float(g["FPP"])) < 1e-6 and abs(tgt.NFPP - float(g["NFPP"])) < 1e-6
    np.testing.assert_allclose(tgt.probs.R_p.values, g["probs/R_p"], rtol=1e-9)


def _compare(g, o):
    assert np.array_equal(g.mask, o.mask)                      # masks: bit-exact
    assert g.n_pass == int(o.mask.sum())
    fin = np.isfinite(o.lnL)
    assert np.array_equal(np.isfinite(g.lnL), fin)
    assert np.array_equal(np.isneginf(g.lnL), np.isneginf(o.lnL))
    np.testing.assert_allclose(g.lnL[fin], o.lnL[fin], rtol=1e-9, atol=0)   # per-draw lnL
    assert abs(g.lnZ - o.lnZ) < 1e-6


@pytest.mark.parametrize("is_host", [False, True])
def test_fused_tp_kernel_against_oracle(gpu_engine, toi465_lc, is_host):
    import _oracle_engine
    t, f, s = toi465_lc
    N = 30000
    rng = np.random.default_rng(5)
    args = dict(rp=rng.uniform(0.5, 20, N), P_orb=3.836169,
                inc=np.degrees(np.arccos(rng.random(N))), ecc=rng.beta(0.867, 3.03, N),
                argp=rng.uniform(0, 360, N), mtot=rng.uniform(0.3, 1.2, N),
                rhost=rng.uniform(0.3, 1.1, N), u1=rng.uniform(0.2, 0.6, N),
                u2=rng.uniform(0.1, 0.3, N), cfr=rng.uniform(0.01, 0.9, N),
                lnprior=np.where(rng.random(N) < 0.1, -np.inf, rng.uniform(-9, 0, N)),
                extra_mask=rng.random(N) < 0.8)
    gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    ora = _oracle_engine.OracleEngine()
    ora.set_lightcurve(t, f, s, 0.00139, 20)
    _compare(gpu_engine.eval_tp(N, **args, companion_is_host=is_host, want_mask=True),
             ora.eval_tp(N, **args, companion_is_host=is_host))


@pytest.mark.parametrize("is_host", [False, True])
def test_fused_eb_kernel_against_oracle(gpu_engine, kepler10b_lc, is_host):
    import _oracle_engine
    t, f, s = kepler10b_lc
    N = 20000
    rng = np.random.default_rng(6)
    q = rng.uniform(0.1, 1.0, N)
    args = dict(reb=0.1 + 0.9 * q, ebfr=0.3 * q ** 3 + 1e-4, q=q, P_orb=rng.uniform(0.8, 0.9, N),
                inc=np.degrees(np.arccos(rng.random(N))), ecc=rng.random(N) ** 5,
                argp=rng.uniform(0, 360, N), mtot=1.0 + q, rhost=rng.uniform(0.8, 1.1, N),
                u1=0.4, u2=0.26, cfr=rng.uniform(0.01, 0.5, N),
                lnprior=rng.uniform(-9, 0, N), extra_mask=rng.random(N) < 0.9)
    gpu_engine.set_lightcurve(t, f, s, 0.0204, 20)
    ora = _oracle_engine.OracleEngine()
    ora.set_lightcurve(t, f, s, 0.0204, 20)
    g0, g1 = gpu_engine.eval_eb(N, **args, companion_is_host=is_host, want_mask=True)
    o0, o1 = ora.eval_eb(N, **args, companion_is_host=is_host)
    _compare(g0, o0)
    _compare(g1, o1)
    assert not np.any(g0.mask & g1.mask)        # q < 0.95 and q >= 0.95 are exclusive


def test_empty_scenario(gpu_engine, toi465_lc):
    t, f, s = toi465_lc
    gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    z = np.zeros(0)
    r = gpu_engine.eval_tp(0, z, 3.8, z, z, z, 0.8, 0.8, 0.4, 0.2, 0.0)
    assert r.lnZ == -np.inf and r.n_pass == 0
    # no draw survives the mask (face-on orbits): every lnL is -inf, lnZ is -inf
    N = 1000
    r = gpu_engine.eval_tp(N, np.full(N, 1.0), 3.8, np.full(N, 10.0), np.zeros(N), np.zeros(N),
                           0.8, 0.8, 0.4, 0.2, 0.0, want_mask=True)
    assert r.lnZ == -np.inf and r.n_pass == 0 and not r.mask.any()
    assert np.all(np.isneginf(r.lnL))


CASES = {
    "underflow": (-2000.0 * np.ones(100000), -2000.0),
    "mixed": (np.array([-1001., -1002., -np.inf, -np.inf, -1003., -np.inf, -1004., -np.inf,
                        -1005., -np.inf]), None),
    "denominator": (np.array([-1.0] + [-np.inf] * 9), -1.0 - np.log(10)),
    "all_neginf": (np.full(50, -np.inf), -np.inf),
    "single": (np.array([-3.5]), -3.5),
    "nan_is_zero_weight": (np.array([-10., np.nan, -11., np.nan]), None),
    "posinf": (np.array([-10., np.inf, -11.]), np.inf),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_log_mean_exp_known_answers(gpu_engine, name):
    """Cases of the reference's tests/test_log_mean_exp.py on the fused reduction."""
    from triceratops_b200._numerics import _log_mean_exp
    lnw, want = CASES[name]
    if want is None:
        want = _log_mean_exp(lnw, N_total=lnw.size)
    got, _ = gpu_engine.log_mean_exp(lnw)
    assert got == want or abs(got - want) < 1e-10


def test_gpu_log_mean_exp_random(gpu_engine):
    from triceratops_b200._numerics import _log_mean_exp
    rng = np.random.default_rng(0)
    lnw = rng.uniform(-5000, -400, 1_000_003)
    lnw[rng.random(lnw.size) < 0.9] = -np.inf
    got, _ = gpu_engine.log_mean_exp(lnw)
    assert abs(got - _log_mean_exp(lnw, N_total=lnw.size)) < 1e-9


def _host_best(lnL, k):
    """Head of a stable sort of -lnL over the finite entries (ties by ascending index)."""
    idx = np.flatnonzero(np.isfinite(lnL))
    order = np.lexsort((idx, -lnL[idx]))
    return idx[order][:k]


@pytest.mark.parametrize("k", [1, 100, 1000])
def test_device_topk_equals_host_sort(gpu_engine, toi465_lc, k):
    t, f, s = toi465_lc
    gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    N = 200_000
    rng = np.random.default_rng(8)
    args = (N, rng.uniform(0.5, 20, N), 3.836169, np.degrees(np.arccos(rng.random(N))),
            rng.beta(0.867, 3.03, N), rng.uniform(0, 360, N), 0.811, 0.84738, 0.43, 0.2, 0.0)
    r = gpu_engine.eval_tp(*args, want_lnL=True, n_best=k)
    want = _host_best(r.lnL, k)
    assert r.n_evaluated == np.isfinite(r.lnL).sum()
    assert np.array_equal(r.top_idx, want)
    assert np.array_equal(r.top_lnL, r.lnL[want])
    assert np.array_equal(gpu_engine.fetch_lnl(0, N), r.lnL)


def test_device_topk_ties_and_short_lists(gpu_engine, toi465_lc):
    """Blocks of identical draws give bit-identical lnL: ties must come out in index order; with
    fewer finite draws than requested the list is simply shorter."""
    t, f, s = toi465_lc
    gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    rng = np.random.default_rng(9)
    base = 37
    rp = np.repeat(rng.uniform(1, 15, base), 300)
    inc = np.repeat(rng.uniform(87.5, 90, base), 300)
    ecc = np.repeat(rng.uniform(0, 0.3, base), 300)
    argp = np.repeat(rng.uniform(0, 360, base), 300)
    N = rp.size
    r = gpu_engine.eval_tp(N, rp, 3.836169, inc, ecc, argp, 0.811, 0.84738, 0.43, 0.2, 0.0,
                           want_lnL=True, n_best=100)
    assert len(np.unique(r.lnL[np.isfinite(r.lnL)])) <= base
    assert np.array_equal(r.top_idx, _host_best(r.lnL, 100))
    # only 5 transiting draws among face-on ones
    inc2 = np.full(N, 5.0)
    inc2[[3, 77, 1000, 5000, 11000]] = 89.5
    r = gpu_engine.eval_tp(N, rp, 3.836169, inc2, np.zeros(N), argp, 0.811, 0.84738, 0.43, 0.2,
                           0.0, want_lnL=True, n_best=100)
    assert r.n_evaluated == 5 and len(r.top_idx) == 5
    assert np.array_equal(r.top_idx, _host_best(r.lnL, 100))


def test_eb_branches_have_their_own_best_lists(gpu_engine, kepler10b_lc):
    t, f, s = kepler10b_lc
    gpu_engine.set_lightcurve(t, f, s, 0.0204, 20)
    N = 60_000
    rng = np.random.default_rng(10)
    q = rng.uniform(0.1, 1.0, N)
    r0, r1 = gpu_engine.eval_eb(N, 0.1 + 0.9 * q, 0.3 * q ** 3 + 1e-4, q, 0.837,
                                np.degrees(np.arccos(rng.random(N))), rng.random(N) ** 5,
                                rng.uniform(0, 360, N), 1.0 + q, 1.0, 0.4, 0.26, 0.05,
                                want_lnL=True, n_best=100)
    for r in (r0, r1):
        assert np.array_equal(r.top_idx, _host_best(r.lnL, 100))
    assert np.all(q[r0.top_idx] < 0.95) and np.all(q[r1.top_idx] >= 0.95)


@pytest.mark.parametrize("name", ["NTPu", "NEBu", "NTPe", "NEBe"])
def test_unknown_and_evolved_nearby_star_functions(name, gpu_engine, golden, toi465_lc, trilegal_file):
    """lnZ_NTP_unknown / NEB_unknown / NTP_evolved / NEB_evolved (marginal_likelihoods.py:2365-3178)."""
    g = golden("lnz_nearby.npz")
    calls = nearby_calls(int(g["N"]), trilegal_file, toi465_lc)
    np.random.seed(int(g["seed"]))
    check_against_golden(name, calls[name](ml), g, lnz_atol=1e-6, arr_rtol=1e-9)


def test_unknown_star_without_similar_trilegal_stars(gpu_engine, toi465_lc, trilegal_file):
    t, f, s = toi465_lc
    res = ml.lnZ_NTP_unknown(t, f, s, 3.8, 40.0, trilegal_file, 100, True)
    assert res["lnZ"] == -np.inf and "b" not in res and res["M_s"] == 0
    res = ml.lnZ_NEB_unknown(t, f, s, 3.8, 40.0, trilegal_file, 100, True)
    assert isinstance(res, dict) and res["lnZ"] == -np.inf and res["b"] == 0
