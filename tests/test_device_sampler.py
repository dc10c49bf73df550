"""Opt-in device sampler (device_priors.py / device_sampler.py): the draws come from a different
random stream than numpy's, so equivalence with the host mode is statistical.

CPU part: every column each scenario hands to the engine is captured in both modes and compared
with a two-sample Kolmogorov-Smirnov test (fixed seeds; all ten scenarios, with and without a
contrast curve), and the deterministic transforms are checked against the host ones on the same
deviates.  GPU part: evidences of the two modes agree within their Monte-Carlo scatter."""
import numpy as np
import pytest
import torch
from scipy import stats

import triceratops_b200
import triceratops_b200.marginal_likelihoods as ml
from conftest import TOI465, lnz_calls
from triceratops_b200 import _dispatch, device_priors as dp, funcs, priors
from triceratops_b200._ldc import grid_for


class _Capture:
    """Stands where the engine stands and keeps the columns of every call."""
    device = -1
    torch_device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def set_lightcurve(self, *a):
        pass

    class _R:
        pass

    def _res(self, N):
        r = self._R()
        r.lnZ, r.m, r.s, r.n_finite, r.n_posinf, r.n_pass = -1.0, -1.0, 1.0, 1, 0, 0
        r.lnL = np.zeros(N)
        r.N, r.top_idx, r.top_lnL, r.n_evaluated, r.branch = N, np.arange(min(N, 100)), \
            np.zeros(min(N, 100)), N, 0
        return r

    @staticmethod
    def _np(v):
        return v.detach().cpu().numpy() if torch.is_tensor(v) else v

    def eval_tp(self, N, rp, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr, lnprior=None,
                extra_mask=None, companion_is_host=False, **kw):
        self.calls.append(dict(rp=rp, P_orb=P_orb, inc=inc, ecc=ecc, argp=argp, mtot=mtot,
                               rhost=rhost, u1=u1, u2=u2, cfr=cfr, lnprior=lnprior,
                               extra_mask=extra_mask, is_host=companion_is_host))
        return self._res(N)

    def eval_eb(self, N, reb, ebfr, q, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr,
                lnprior=None, extra_mask=None, companion_is_host=False, **kw):
        self.calls.append(dict(reb=reb, ebfr=ebfr, q=q, P_orb=P_orb, inc=inc, ecc=ecc, argp=argp,
                               mtot=mtot, rhost=rhost, u1=u1, u2=u2, cfr=cfr, lnprior=lnprior,
                               extra_mask=extra_mask, is_host=companion_is_host))
        return self._res(N), self._res(N)

    def eval_tp_tensors(self, N, cols, extra_mask=None, companion_is_host=False, n_best=100):
        self.calls.append(dict({k: self._np(v) for k, v in cols.items()},
                               extra_mask=self._np(extra_mask), is_host=companion_is_host))
        return self._res(N)

    def eval_eb_tensors(self, N, cols, extra_mask=None, companion_is_host=False, n_best=100):
        self.calls.append(dict({k: self._np(v) for k, v in cols.items()},
                               extra_mask=self._np(extra_mask), is_host=companion_is_host))
        return self._res(N), self._res(N)


@pytest.fixture()
def capture():
    cap = _Capture()
    saved = _dispatch.get_engine
    _dispatch.get_engine = lambda: cap
    yield cap
    _dispatch.get_engine = saved
    triceratops_b200.set_sampler("host")


NAMES = ["TTP", "TEB", "PTP", "PTPcc", "PEB", "STP", "STPcc", "SEB", "SEBcc", "DTP", "DTPcc",
         "DEB", "BTP", "BTPcc", "BEB", "BEBcc"]


@pytest.mark.parametrize("name", NAMES)
def test_columns_have_the_same_distributions_in_both_modes(name, capture, toi465_lc,
                                                           trilegal_file, contrast_file):
    N = 60_000
    calls = lnz_calls(TOI465, N, trilegal_file, contrast_file, toi465_lc)
    triceratops_b200.set_sampler("host")
    np.random.seed(101)
    calls[name](ml)
    host = capture.calls.pop()
    triceratops_b200.set_sampler("device", seed=202)
    calls[name](ml)
    dev = capture.calls.pop()
    assert host["is_host"] == dev["is_host"]
    assert set(host) == set(dev)
    for key in host:
        if key == "is_host":
            continue
        h, d = host[key], dev[key]
        if h is None or d is None:
            if key == "extra_mask":       # "all true" may be passed as None
                assert (h is None or np.all(h)) and (d is None or np.all(d)), (name, key)
            else:
                assert h is None and d is None, (name, key)
            continue
        h, d = np.asarray(h, float).ravel(), np.asarray(d, float).ravel()
        if h.size == 1 or d.size == 1:
            assert h.size == d.size == 1 and np.isclose(h[0], d[0], rtol=1e-12), (name, key)
            continue
        # same share of -inf / excluded entries (binomial tolerance), same body
        for bad in (np.isneginf, lambda v: ~np.isfinite(v)):
            fh, fd = bad(h).mean(), bad(d).mean()
            assert abs(fh - fd) < 5 * np.sqrt(max(fh, 1e-4) / N) + 1e-3, (name, key, fh, fd)
        h, d = h[np.isfinite(h)], d[np.isfinite(d)]
        p = stats.ks_2samp(h, d).pvalue
        assert p > 1e-4, (name, key, p)


def test_transforms_agree_with_host_on_the_same_deviates(contrast_file):
    rng = np.random.default_rng(0)
    n = 50_000
    x = rng.random(n)
    xt = torch.from_numpy(x.copy())
    Ms = rng.uniform(0.1, 1.5, n)
    close = lambda a, b: np.testing.assert_allclose(np.asarray(b), a, rtol=1e-11)  # noqa: E731
    close(priors.sample_rp(x.copy(), Ms, False), dp.sample_rp(xt, torch.from_numpy(Ms), False))
    close(priors.sample_inc(x.copy()), dp.sample_inc(xt))
    for M in (1.3, 0.811, 0.25, 0.08):
        close(priors.sample_q(x.copy(), M), dp.sample_q(xt, M))
        close(priors.sample_q_companion(x.copy(), M), dp.sample_q_companion(xt, M))
    m = rng.uniform(0.05, 3, n)
    mt = torch.from_numpy(m)
    a = funcs.stellar_relations(m, np.full(n, 0.9), np.full(n, 5000.))
    b = dp.stellar_relations(mt, 0.9, 5000.)
    close(a[0], b[0])
    close(a[1], b[1])
    for filt in ("TESS", "J", "H", "K"):
        close(funcs.flux_relation(m, filt), dp.flux_relation(mt, filt))
    sep, con = funcs.file_to_contrast_curve(contrast_file)
    con = np.maximum.accumulate(con) + np.arange(con.size) * 1e-9     # a monotonic curve
    dm = rng.uniform(0, 12, n)
    with np.errstate(divide="ignore"):
        for M in (1.3, 0.811):
            for host_fn, dev_fn in ((priors.lnprior_bound_TP, dp.lnprior_bound_TP),
                                    (priors.lnprior_bound_EB, dp.lnprior_bound_EB)):
                want = host_fn(M, 8.16, dm, sep, con)
                got = dev_fn(M, 8.16, torch.from_numpy(dm), torch.from_numpy(sep),
                             torch.from_numpy(con)).numpy()
                assert np.array_equal(np.isfinite(want), np.isfinite(got))
                fin = np.isfinite(want)
                np.testing.assert_allclose(got[fin], want[fin], rtol=1e-10)
    T, lg = rng.uniform(2800, 9900, n), rng.uniform(3, 5.6, n)
    a = grid_for("TESS").at_Z_rounded(0.0, T, lg, 10000)
    b = dp.ldc_at_Z_rounded(grid_for("TESS"), 0.0, torch.from_numpy(T), torch.from_numpy(lg),
                            10000)
    assert np.array_equal(a[0], b[0].numpy()) and np.array_equal(a[1], b[1].numpy())


def test_eccentricity_samplers_follow_their_laws():
    torch.manual_seed(3)
    n = 200_000
    assert stats.kstest(dp.sample_ecc(n, True, 3.0, "cpu").numpy(), "beta",
                        args=(0.867, 3.03)).pvalue > 1e-3
    assert stats.kstest(dp.sample_ecc(n, False, 3.0, "cpu").numpy(), "powerlaw",
                        args=(0.2,)).pvalue > 1e-3
    assert stats.kstest(dp.sample_ecc(n, False, 20.0, "cpu").numpy(), "powerlaw",
                        args=(0.6,)).pvalue > 1e-3


def test_result_tables_keep_the_reference_layout(oracle_engine, toi465_lc):
    t, f, s = toi465_lc
    triceratops_b200.set_sampler("device", seed=5)
    try:
        res, twin = ml.lnZ_TEB(t, f, s, 3.836169, 0.811, 0.84738, 4936.0, 0.0, 1500, True)
    finally:
        triceratops_b200.set_sampler("host")
    for r in (res, twin):
        assert set(r) == {"M_s", "R_s", "u1", "u2", "P_orb", "inc", "b", "R_p", "ecc", "argp",
                          "M_EB", "R_EB", "fluxratio_EB", "fluxratio_comp", "lnZ"}
        assert all(len(r[k]) == 100 for k in r if k != "lnZ")
    assert np.all(twin["P_orb"] == 2 * 3.836169)


@pytest.mark.gpu
def test_device_and_host_evidences_agree_within_monte_carlo_scatter(gpu_engine, toi465_lc,
                                                                    trilegal_file, contrast_file):
    """lnZ is an average over N draws: with N = 4e5 and 6 independent repeats per mode the two
    modes must agree within the scatter of the repeats."""
    N, reps = 400_000, 6
    calls = lnz_calls(TOI465, N, trilegal_file, contrast_file, toi465_lc)
    try:
        for name in ("TTP", "PTPcc", "STP", "DTP", "BTP", "TEB", "SEB", "BEBcc"):
            host, dev = [], []
            for r in range(reps):
                triceratops_b200.set_sampler("host")
                np.random.seed(1000 + r)
                h = calls[name](ml)
                triceratops_b200.set_sampler("device", seed=2000 + r)
                d = calls[name](ml)
                host.append([x["lnZ"] for x in (h if isinstance(h, tuple) else (h,))])
                dev.append([x["lnZ"] for x in (d if isinstance(d, tuple) else (d,))])
            host, dev = np.array(host), np.array(dev)
            for b in range(host.shape[1]):
                hb, db = host[:, b], dev[:, b]
                if not np.all(np.isfinite(hb)) or not np.all(np.isfinite(db)):
                    continue
                # compare in evidence space: the mean of Z over repeats
                zh = np.exp(hb - hb.max())
                zd = np.exp(db - hb.max())
                err = np.sqrt(zh.var(ddof=1) / reps + zd.var(ddof=1) / reps)
                assert abs(zh.mean() - zd.mean()) < 5 * err + 0.05 * zh.mean(), (name, b, hb, db)
    finally:
        triceratops_b200.set_sampler("host")


@pytest.mark.gpu
def test_device_mode_calc_probs_runs_and_is_reproducible(gpu_engine, toi465_lc, trilegal_file,
                                                         contrast_file):
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    t, f, s = toi465_lc
    stars = synth.stars_table(270380593, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"],
                              TOI465["M"], TOI465["R"], TOI465["Teff"], TOI465["plx"])
    out = []
    try:
        for _ in range(2):
            triceratops_b200.set_sampler("device", seed=77)
            tgt = target(270380593, stars=stars, trilegal_fname=trilegal_file)
            tgt.calc_probs(t, f, s, TOI465["P"], contrast_curve_file=contrast_file, filt="K",
                           N=200_000, parallel=True, verbose=0)
            out.append(tgt.lnZ.copy())
            assert len(tgt.probs) == 18 and abs(tgt.probs.prob.sum() - 1) < 1e-9
    finally:
        triceratops_b200.set_sampler("host")
    assert np.array_equal(out[0], out[1])      # same seed, same streams, same answer


@pytest.mark.gpu
def test_splev_kernel_matches_scipy(gpu_engine):
    """tri_dev_splev (one launch per spline evaluation in device mode) against scipy's FITPACK,
    inside the knot range and extrapolating beyond both ends."""
    import torch
    from scipy.interpolate import InterpolatedUnivariateSpline, splev as sp_splev
    from triceratops_b200 import device_priors as dp, funcs
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(0.05, 2.6, 200_000), [0.0, 0.1, 0.63, 2.0, 3.5, -1.0]])
    xd = torch.as_tensor(x, device="cuda")
    splines = [funcs._hot_R, funcs._cool_R, funcs._hot_T, funcs._cool_T]
    splines += list(funcs._FLUX_SPLINES.values())
    for k in (1, 2, 4, 5):     # the stellar relations are cubic; cover the other degrees too
        xs = np.linspace(0, 3, 15)
        splines.append(InterpolatedUnivariateSpline(xs, np.sin(xs) + xs ** 2, k=k))
    for spl in splines:
        want = sp_splev(x, spl._eval_args, ext=0)
        got = dp.splev(spl, xd).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13 * np.abs(want).max())
