import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


TOI465 = dict(P=3.836169, M=0.811, R=0.84738, Teff=4936.0, plx=8.16366,
              T=10.7307, J=9.906, H=9.473, K=9.339)
KEP10 = dict(P=0.837, M=1.017, R=1.08974, Teff=5706.0, plx=5.36185,
             T=10.4, J=9.889, H=9.563, K=9.496)


def load_lc(name):
    lc = np.loadtxt(os.path.join(GOLD, name), delimiter=",")
    return lc[:, 0].copy(), lc[:, 1].copy(), float(np.mean(lc[:, 2]))


@pytest.fixture(scope="session")
def toi465_lc():
    return load_lc("TOI465_01_lightcurve.csv")


@pytest.fixture(scope="session")
def kepler10b_lc():
    return load_lc("Kepler10b_lightcurve.csv")


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        return np.load(os.path.join(GOLD, name), allow_pickle=False)
    return _load


@pytest.fixture(scope="session")
def trilegal_file():
    return os.path.join(GOLD, "trilegal_synth.csv")


@pytest.fixture(scope="session")
def contrast_file():
    return os.path.join(GOLD, "TOI465_01_contrastcurve.csv")


@pytest.fixture()
def oracle_engine():
    """Routes the host layer to the CPU oracle stand-in for the duration of a test."""
    import _oracle_engine
    eng = _oracle_engine.install()
    yield eng
    _oracle_engine.uninstall()


@pytest.fixture(scope="session")
def gpu_engine():
    """The real CUDA engine; fails (does not skip) if the extension cannot run."""
    from triceratops_b200.engine import get_engine
    return get_engine()


def lnz_calls(star, N, tri, cc, lc, mission="TESS", exptime=0.00139):
    """The scenario calls of oracle/gen_golden.py, against any module exposing lnZ_*."""
    t, f, s = lc
    base = (t, f, s, star["P"], star["M"], star["R"], star["Teff"])
    tail = (N, True, mission, False, exptime, 20)
    mags = (star.get("T"), star.get("J"), star.get("H"), star.get("K"))
    calls = {
        "TTP": lambda m: m.lnZ_TTP(*base, 0.0, *tail),
        "TEB": lambda m: m.lnZ_TEB(*base, 0.0, *tail),
    }
    if tri is None:
        return calls
    calls.update({
        "PTP": lambda m: m.lnZ_PTP(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "PTPcc": lambda m: m.lnZ_PTP(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "PEB": lambda m: m.lnZ_PEB(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "PEBcc": lambda m: m.lnZ_PEB(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "STP": lambda m: m.lnZ_STP(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "STPcc": lambda m: m.lnZ_STP(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "SEB": lambda m: m.lnZ_SEB(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "SEBcc": lambda m: m.lnZ_SEB(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "DTP": lambda m: m.lnZ_DTP(*base, 0.0, *mags, tri, None, "TESS", *tail),
        "DTPcc": lambda m: m.lnZ_DTP(*base, 0.0, *mags, tri, cc, "K", *tail),
        "DEB": lambda m: m.lnZ_DEB(*base, 0.0, *mags, tri, None, "TESS", *tail),
        "DEBcc": lambda m: m.lnZ_DEB(*base, 0.0, *mags, tri, cc, "J", *tail),
        "BTP": lambda m: m.lnZ_BTP(*base, *mags, tri, None, "TESS", *tail),
        "BTPcc": lambda m: m.lnZ_BTP(*base, *mags, tri, cc, "H", *tail),
        "BEB": lambda m: m.lnZ_BEB(*base, *mags, tri, None, "TESS", *tail),
        "BEBcc": lambda m: m.lnZ_BEB(*base, *mags, tri, cc, "K", *tail),
    })
    return calls


def scalar_calls(star, N, tri, cc, lc, parallel=False):
    """The calls of oracle/gen_golden.py's lnz_scalar.npz (the reference's parallel=False
    loops); parallel=True gives the vectorised semantics on the same draws."""
    t, f, s = lc
    base = (t, f, s, star["P"], star["M"], star["R"], star["Teff"])
    tail = (N, parallel, "TESS", False, 0.00139, 20)
    mags = (star.get("T"), star.get("J"), star.get("H"), star.get("K"))
    return {
        "TTP": lambda m: m.lnZ_TTP(*base, 0.0, *tail),
        "TEB": lambda m: m.lnZ_TEB(*base, 0.0, *tail),
        "PTP": lambda m: m.lnZ_PTP(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "PEBcc": lambda m: m.lnZ_PEB(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "STPcc": lambda m: m.lnZ_STP(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "SEB": lambda m: m.lnZ_SEB(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "DTP": lambda m: m.lnZ_DTP(*base, 0.0, *mags, tri, None, "TESS", *tail),
        "DEBcc": lambda m: m.lnZ_DEB(*base, 0.0, *mags, tri, cc, "J", *tail),
        "BTPcc": lambda m: m.lnZ_BTP(*base, *mags, tri, cc, "H", *tail),
        "BEB": lambda m: m.lnZ_BEB(*base, *mags, tri, None, "TESS", *tail),
    }


SCALAR_NAMES = ["TTP", "TEB", "PTP", "PEBcc", "STPcc", "SEB", "DTP", "DEBcc", "BTPcc", "BEB"]


def scalar_star(gold, tag):
    if tag == "toi465":
        return TOI465
    return {k: float(gold["tight_star/" + k]) for k in ("P", "M", "R", "Teff", "plx", "T", "J",
                                                        "H", "K")}


class _Prefixed:
    """View of a fixture under a key prefix (lnz_scalar.npz holds two stars)."""

    def __init__(self, gold, prefix):
        self.gold, self.prefix = gold, prefix

    def __getitem__(self, key):
        return self.gold[self.prefix + key]


def nearby_calls(N, tri, lc):
    """The unknown / evolved nearby-star calls of oracle/gen_golden.py."""
    t, f, s = lc
    P = TOI465["P"]
    return {
        "NTPu": lambda m: m.lnZ_NTP_unknown(t, f, s, P, 13.0, tri, N, True),
        "NEBu": lambda m: m.lnZ_NEB_unknown(t, f, s, P, 13.0, tri, N, True),
        "NTPe": lambda m: m.lnZ_NTP_evolved(t, f, s, P, 2.5, 5100.0, 0.0, N, True),
        "NEBe": lambda m: m.lnZ_NEB_evolved(t, f, s, P, 2.5, 5100.0, 0.0, N, True),
    }


RESULT_KEYS = ("M_s", "R_s", "u1", "u2", "P_orb", "inc", "b", "R_p", "ecc", "argp", "M_EB",
               "R_EB", "fluxratio_EB", "fluxratio_comp")


def check_against_golden(name, res, gold, lnz_atol=1e-6, arr_rtol=1e-9):
    """lnZ within tolerance (north_star: 1e-6); best-draw tables equal up to rounding.

    Draws with bit-identical lnL (e.g. every draw whose eclipse misses the observed window) are
    ordered arbitrarily by the reference's unstable argsort (marginal_likelihoods.py:153) and by
    index here, so a table that differs row-for-row must still match as a multiset of rows.
    """
    branches = res if isinstance(res, tuple) else (res,)
    for b, r in enumerate(branches):
        g_lnz = float(gold["%s/%d/lnZ" % (name, b)])
        if np.isfinite(g_lnz):
            # 1e-6 absolute (north_star) plus the 1e-9 relative allowed on each draw's lnL,
            # which dominates when |lnZ| is huge (a scenario that fits nothing)
            assert abs(r["lnZ"] - g_lnz) <= lnz_atol + 1e-9 * abs(g_lnz), \
                (name, b, r["lnZ"], g_lnz)
        else:
            assert r["lnZ"] == g_lnz, (name, b, r["lnZ"], g_lnz)
        mine = np.stack([np.asarray(r[k], float) for k in RESULT_KEYS], axis=1)
        want = np.stack([gold["%s/%d/%s" % (name, b, k)] for k in RESULT_KEYS], axis=1)
        n_eval = getattr(r, "n_evaluated", None)
        if n_eval is not None and n_eval < len(mine):
            # rows beyond the evaluated draws carry zero weight; the reference fills them in
            # the arbitrary order of numpy's argsort over -inf entries
            mine, want = mine[:n_eval], want[:n_eval]
        if not np.allclose(mine, want, rtol=arr_rtol, atol=0, equal_nan=True):
            mine = mine[np.lexsort(mine.T[::-1])]
            want = want[np.lexsort(want.T[::-1])]
        np.testing.assert_allclose(mine, want, rtol=arr_rtol, atol=0,
                                   err_msg="%s branch %d best-draw table" % (name, b))


def draw_tp_columns(N, seed, star=None):
    """Prior draws of a TP-type scenario as keyword columns of Engine.eval_tp / submit_tp."""
    from triceratops_b200 import priors
    star = star or TOI465
    rng_state = np.random.get_state()
    np.random.seed(seed)
    try:
        rp = priors.sample_rp(np.random.rand(N), np.full(N, star["M"]), False)
        inc = priors.sample_inc(np.random.rand(N))
        ecc = priors.sample_ecc(np.random.rand(N), True, star["P"])
        argp = priors.sample_w(np.random.rand(N))
    finally:
        np.random.set_state(rng_state)
    return dict(rp=rp, P_orb=star["P"], inc=inc, ecc=ecc, argp=argp, mtot=star["M"],
                rhost=star["R"], u1=0.4338, u2=0.2008, cfr=0.0)


def draw_eb_columns(N, seed, star=None):
    """Prior draws of an EB-type scenario as keyword columns of Engine.eval_eb / submit_eb."""
    from triceratops_b200 import funcs, priors
    star = star or TOI465
    rng_state = np.random.get_state()
    np.random.seed(seed)
    try:
        inc = priors.sample_inc(np.random.rand(N))
        q = priors.sample_q(np.random.rand(N), star["M"])
        ecc = priors.sample_ecc(np.random.rand(N), False, star["P"])
        argp = priors.sample_w(np.random.rand(N))
    finally:
        np.random.set_state(rng_state)
    masses = q * star["M"]
    radii, _ = funcs.stellar_relations(masses, np.full(N, star["R"]), np.full(N, star["Teff"]))
    fr = funcs.flux_relation(masses) / (funcs.flux_relation(masses)
                                        + funcs.flux_relation(np.array([star["M"]])))
    return dict(reb=radii, ebfr=fr, q=q, P_orb=star["P"], inc=inc, ecc=ecc, argp=argp,
                mtot=star["M"] + masses, rhost=star["R"], u1=0.4338, u2=0.2008, cfr=0.0)


def calc_probs_small(t, f, s, N=301, seed=77, full=False):
    """A short calc_probs (three target-star scenario calls) used by the multi-rank test."""
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    stars = synth.stars_table(9, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"], TOI465["M"],
                              TOI465["R"], TOI465["Teff"], TOI465["plx"], n_neighbours=0)
    tgt = target(9, stars=stars, trilegal_fname=os.path.join(GOLD, "trilegal_synth.csv"))
    np.random.seed(seed)
    tgt.calc_probs(t, f, s, TOI465["P"], N=N, parallel=True, verbose=0,
                   drop_scenario=["PTP", "PEB", "STP", "SEB", "DEB", "BTP", "BEB"])
    return tgt if full else tgt.lnZ.copy()
