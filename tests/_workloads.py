"""BASELINE.json's configurations as concrete inputs, shared by the
parity tests and bench.py, and the recorder that captures the engine calls `target.calc_probs`
makes (so that tests and the bench feed the kernels exactly what calc_probs feeds them).

Test / bench infrastructure: nothing here is imported by the triceratops_b200 package.
"""
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

TOI465 = dict(ID=270380593, P=3.836169, M=0.811, R=0.84738, Teff=4936.0, plx=8.16366,
              T=10.7307, J=9.906, H=9.473, K=9.339)
KEP10 = dict(ID=11904151, P=0.837, M=1.017, R=1.08974, Teff=5706.0, plx=5.36185,
             T=10.4, J=9.889, H=9.563, K=9.496)
SOLAR = dict(ID=1, P=10.0, M=1.0, R=1.0, Teff=5750.0, plx=10.0, T=10.0, J=9.2, H=8.9, K=8.8)

# config number -> (label, star, default draws per scenario, mission, exptime)
CONFIGS = {
    1: ("configs[0]: TOI-465.01 858-stamp light curve, lnZ_TTP + lnZ_TEB only", TOI465,
        100_000, "TESS", 0.00139),
    2: ("configs[1]: TOI-465.01 858-stamp folded light curve, 18 scenario rows (15 target + "
        "NTP/NEB/NEBx2P), contrast curve, nsamples=20", TOI465, 1_000_000, "TESS", 0.00139),
    3: ("configs[2]: Kepler-10b 478-stamp light curve, 29.4-min exposures supersampled 20x, 18 "
        "scenario rows", KEP10, 1_000_000, "Kepler", 0.0204),
    4: ("configs[3]: synthetic 20000-stamp 2-min folded light curve (P = 10 d), 18 scenario rows",
        SOLAR, 10_000_000, "TESS", 0.00139),
}


def _csv_lc(name):
    lc = np.loadtxt(os.path.join(GOLD, name), delimiter=",")
    return lc[:, 0].copy(), lc[:, 1].copy(), float(np.mean(lc[:, 2]))


def config4_lightcurve(model):
    """20 000 two-minute stamps on [-0.5, 0.5] d, injected k = 0.05,
    a/R* = 15, b = 0.3 circular transit (u = 0.4, 0.2; P = 10 d) + N(0, 1e-3) noise.
    `model(t, k, P, a_rs, inc_rad, e, w_rad, u1, u2, exptime, nsamples)` supplies the transit:
    the oracle in the tests, the engine's simulate seam in bench.py."""
    t = np.linspace(-0.5, 0.5, 20000)
    truth = model(t, 0.05, 10.0, 15.0, np.arccos(0.3 / 15.0), 0.0, np.pi / 2, 0.4, 0.2,
                  0.00139, 20)
    return t, truth + np.random.default_rng(1234).normal(0, 1e-3, t.size), 1e-3


def engine_model(t, k, P, a_rs, inc_rad, e, w_rad, u1, u2, exptime, nsamples):
    """The transit model through the package's own simulate seam (GPU)."""
    from triceratops_b200._constants import Rearth, Rsun
    from triceratops_b200.likelihoods import simulate_TP_transit
    assert e == 0.0
    return simulate_TP_transit(t, k * Rsun / Rearth, P, np.degrees(inc_rad), a_rs * Rsun, 1.0,
                               u1, u2, 0.0, 0.0, exptime=exptime, nsamples=nsamples)


def lightcurve(config, model=None):
    if config in (1, 2):
        return _csv_lc("TOI465_01_lightcurve.csv")
    if config == 3:
        return _csv_lc("Kepler10b_lightcurve.csv")
    if model is None:
        from oracle import coracle
        model = coracle.model
    return config4_lightcurve(model)


def make_target(config):
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    star = CONFIGS[config][1]
    stars = synth.stars_table(star["ID"], star["T"], star["J"], star["H"], star["K"], star["M"],
                              star["R"], star["Teff"], star["plx"])
    return target(star["ID"], stars=stars, mission=CONFIGS[config][3],
                  trilegal_fname=os.path.join(GOLD, "trilegal_synth.csv"))


def run_calc_probs(tgt, config, lc, N, seed, **kw):
    """One full `target.calc_probs` of the configuration (numpy seed set first)."""
    t, f, s = lc
    star, exptime = CONFIGS[config][1], CONFIGS[config][4]
    np.random.seed(seed)
    tgt.calc_probs(t, f, s, star["P"],
                   contrast_curve_file=os.path.join(GOLD, "TOI465_01_contrastcurve.csv"),
                   filt="K", N=N, parallel=True, verbose=0, exptime=exptime, nsamples=20, **kw)
    return tgt


class Recorder:
    """Stands where the engine stands while calc_probs' host code runs once, and keeps the
    columns each engine call would receive."""
    device = -1

    def __init__(self):
        self.calls = []
        self.lc = None

    def set_lightcurve(self, time_, flux, sigma, exptime, nsamples):
        self.lc = (np.array(time_, float), np.array(flux, float), float(sigma), float(exptime),
                   int(nsamples))

    class _R:
        pass

    def _dummy(self, N):
        r = self._R()
        r.lnZ, r.m, r.s, r.n_finite, r.n_posinf, r.n_pass = -1.0, -1.0, 1.0, 1, 0, 0
        r.lnL = np.zeros(N)
        return r

    def eval_tp(self, N, rp, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr, lnprior=None,
                extra_mask=None, companion_is_host=False, **kw):
        self.calls.append(dict(kind="tp", N=N, lc=self.lc, is_host=bool(companion_is_host),
                               extra_mask=extra_mask,
                               cols=dict(rp=rp, P_orb=P_orb, inc=inc, ecc=ecc, argp=argp,
                                         mtot=mtot, rhost=rhost, u1=u1, u2=u2, cfr=cfr,
                                         lnprior=lnprior)))
        return self._dummy(N)

    def eval_eb(self, N, reb, ebfr, q, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr,
                lnprior=None, extra_mask=None, companion_is_host=False, **kw):
        self.calls.append(dict(kind="eb", N=N, lc=self.lc, is_host=bool(companion_is_host),
                               extra_mask=extra_mask,
                               cols=dict(reb=reb, ebfr=ebfr, q=q, P_orb=P_orb, inc=inc, ecc=ecc,
                                         argp=argp, mtot=mtot, rhost=rhost, u1=u1, u2=u2,
                                         cfr=cfr, lnprior=lnprior)))
        return self._dummy(N), self._dummy(N)


class patched_engine:
    """Context manager: `_dispatch.get_engine` returns `eng` inside (tests / bench only)."""

    def __init__(self, eng):
        self.eng = eng

    def __enter__(self):
        from triceratops_b200 import _dispatch
        self._saved = _dispatch.get_engine
        _dispatch.get_engine = lambda: self.eng
        return self.eng

    def __exit__(self, *exc):
        from triceratops_b200 import _dispatch
        _dispatch.get_engine = self._saved


def record_calls(config, N, seed, lc):
    """Run calc_probs' host side once (prior draws, stellar relations, priors) and record the
    engine calls of the configuration: 12 for the 18-row table (6 TP-type, 6 EB-type), 2 for
    config 1.  Returns (calls, seconds of host work)."""
    tgt = make_target(config)
    rec = Recorder()
    t0 = time.perf_counter()
    with patched_engine(rec):
        kw = {}
        if config == 1:
            kw["drop_scenario"] = ["PTP", "PEB", "STP", "SEB", "DTP", "DEB", "BTP", "BEB"]
            tgt.stars = tgt.stars.iloc[:1].copy()      # the single target star
        run_calc_probs(tgt, config, lc, N, seed, **kw)
    host_s = time.perf_counter() - t0
    assert len(rec.calls) == (2 if config == 1 else 12), len(rec.calls)
    return rec.calls, host_s
