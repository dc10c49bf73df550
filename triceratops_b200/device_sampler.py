"""The ten `lnZ_*` scenarios with the prior draws generated on the device (opt-in).

Same scenario definitions as `marginal_likelihoods.py` (reference
marginal_likelihoods.py:39-2362) -- same priors, same derived quantities, same wiring into the
two kernels -- but every per-draw column is produced in HBM by `device_priors.py` and handed to
`tri_eval_tp_dev` / `tri_eval_eb_dev` by pointer.  The results are statistically, not bitwise,
equivalent to the host-sampler mode (different random streams); see device_priors.py.

Enable with `triceratops_b200.set_sampler("device", seed=...)`.  Under a process group every
rank draws its own N/G draws (independent streams), so nothing but the evidence records and
best-draw candidates is exchanged.
"""
import math

import numpy as np
import torch
from pandas import read_csv

from . import _dispatch
from . import device_priors as dp
from ._constants import G, Msun, Rsun, pi
from ._ldc import grid_for
from .funcs import file_to_contrast_curve, trilegal_results

N_SAMPLES = 100
_state = {"seed": None, "calls": 0}


def seed(value):
    _state["seed"] = value
    _state["calls"] = 0


def _device():
    """The torch device of the engine's GPU (tests inject an engine whose `torch_device` says
    otherwise; the CUDA engine has no such attribute)."""
    eng = _dispatch.get_engine()
    forced = getattr(eng, "torch_device", None)
    return forced if forced is not None else torch.device("cuda", eng.device)


def _begin(time, flux, sigma, exptime, nsamples, N):
    """Upload the light curve, seed this call's stream, return (device, local draw count)."""
    eng = _dispatch.get_engine()
    _dispatch.use_lightcurve(time, flux, sigma, exptime, nsamples)   # applied with the submission
    dev = _device()
    lo, hi = _dispatch.shard_bounds(N)
    d = _dispatch._dist()
    if _state["seed"] is None and d is not None:
        # unseeded run under a process group: torch's generators start from the same state in
        # every process, so the ranks would all draw the SAME N/G samples.  One base seed is
        # agreed on (rank 0's entropy) and every rank derives its own stream from it below.
        box = [int(torch.seed()) % (2 ** 62)]
        d.broadcast_object_list(box, src=0)
        _state["seed"], _state["auto_seed"] = box[0], True
    if _state["seed"] is not None:
        rank = d.get_rank() if d is not None else 0
        s = (int(_state["seed"]) * 1000003 + _state["calls"] * 7919 + rank) % (2 ** 63 - 1)
        torch.manual_seed(s)
        if dev.type == "cuda":
            torch.cuda.manual_seed(s)
    _state["calls"] += 1
    _state["bounds"], _state["N_total"] = (lo, hi), int(N)
    return eng, dev, hi - lo


def _periods(P_orb, n, dev):
    if type(P_orb) not in [float, int]:
        P = P_orb[0] + (P_orb[-1] - P_orb[0]) * dp.rand(n, dev)
        return P, float(0.5 * (P_orb[0] + P_orb[-1]))
    return float(P_orb), float(P_orb)


def _logg(M, R):
    return math.log10(G * (M * Msun) / (R * Rsun) ** 2)


def _draw_planet(n, host_masses, flatpriors, P_mean, dev):
    rps = dp.sample_rp(dp.rand(n, dev), host_masses, flatpriors)
    incs = dp.sample_inc(dp.rand(n, dev))
    eccs = dp.sample_ecc(n, True, P_mean, dev)
    argps = dp.sample_w(dp.rand(n, dev))
    return rps, incs, eccs, argps


def _draw_binary(n, M_s, P_mean, dev):
    incs = dp.sample_inc(dp.rand(n, dev))
    qs = dp.sample_q(dp.rand(n, dev), M_s)
    eccs = dp.sample_ecc(n, False, P_mean, dev)
    argps = dp.sample_w(dp.rand(n, dev))
    return incs, qs, eccs, argps


def _companion_q(n, M_s, molusc_file, dev):
    if molusc_file is None:
        return dp.sample_q_companion(dp.rand(n, dev), M_s)
    df = read_csv(molusc_file)
    sma = df["semi-major axis(AU)"].values
    e = df["eccentricity"].values
    q = np.array(df[sma * (1 - e) > 10]["mass ratio"].values, dtype=float)
    q[q < 0.1 / M_s] = 0.1 / M_s
    # the table is padded to the TOTAL draw count and this rank takes its slice of it, as the
    # host sampler does (marginal_likelihoods._companion_q): rows are neither dropped nor
    # counted once per rank
    lo, hi = _state["bounds"]
    N = _state["N_total"]
    q = q[:N]
    return dp._t(np.pad(q, (0, N - len(q)))[lo:hi], dev)


def _fluxratio(masses, M_s, filt="TESS"):
    f = dp.flux_relation(masses, filt)
    return f / (f + dp.flux_relation_scalar(M_s, filt))


def _bound_prior(prior_fn, M_s, plx, n, dev, molusc_file, contrast_curve_file, fr_tess, fr_cc_fn):
    if molusc_file is not None:
        return None
    if contrast_curve_file is None:
        dm = 2.5 * torch.log10(fr_tess)
        sep, con = dp._t([2.2], dev), dp._t([1.0], dev)
    else:
        dm = 2.5 * torch.log10(fr_cc_fn())
        s_, c_ = file_to_contrast_curve(contrast_curve_file)
        sep, con = dp._t(s_, dev), dp._t(c_, dev)
    return dp.clip_prior(prior_fn(M_s, plx, torch.abs(dm), sep, con), dm)


class _Background:
    def __init__(self, trilegal_fname, Tmag, Jmag, Hmag, Kmag, dev):
        (Tm, masses, loggs, Teffs, Zs, Jm, Hm, Km) = trilegal_results(trilegal_fname, Tmag)
        self.N_comp = Tm.shape[0]
        self.delta = {"T": Tmag - Tm, "J": Jmag - Jm, "H": Hmag - Hm, "K": Kmag - Km}
        self.h_masses, self.h_loggs, self.h_Teffs, self.h_Zs = masses, loggs, Teffs, Zs
        self.dev = dev
        self.masses, self.loggs, self.Teffs = (dp._t(masses, dev), dp._t(loggs, dev),
                                               dp._t(Teffs, dev))
        self.fluxratios = dp._t(self._fr("T"), dev)

    def _fr(self, band):
        d = self.delta[band]
        return 10 ** (d / 2.5) / (1 + 10 ** (d / 2.5))

    def band_key(self, filt):
        return filt if filt in ("J", "H", "K") else "T"

    def dmag(self, filt):
        return dp._t(self.delta[self.band_key(filt)], self.dev)

    def fluxratios_in(self, filt):
        return dp._t(self._fr(self.band_key(filt)), self.dev)

    def radii(self):
        return dp._t(np.sqrt(G * self.h_masses * Msun / 10 ** self.h_loggs) / Rsun, self.dev)

    def ldc(self, mission):
        u1, u2 = grid_for(mission).nearest_each(self.h_Teffs, self.h_loggs, self.h_Zs)
        return dp._t(u1, self.dev), dp._t(u2, self.dev)


def _background_prior(bg, n, dev, contrast_curve_file, dmag_tess, dmag_cc):
    if contrast_curve_file is None:
        c = math.log((bg.N_comp / 0.1) * (1 / 3600) ** 2 * 2.2 ** 2)
        return dp.clip_prior(torch.full((n,), c, dtype=dp.F64, device=dev), dmag_tess)
    s_, c_ = file_to_contrast_curve(contrast_curve_file)
    lnprior = dp.lnprior_background(bg.N_comp, torch.abs(dmag_cc), dp._t(s_, dev), dp._t(c_, dev))
    return dp.clip_prior(lnprior, dmag_cc)


# ------------------------------------------------------------------------------- result tables
def _take_rows(columns, idx, dev):
    """Rows `idx` of every column (device tensors or scalars) as numpy arrays, with a single
    device-to-host copy for all tensor columns."""
    n = len(idx)
    tens = [k for k, v in columns.items() if torch.is_tensor(v)]
    out = {}
    if tens and n:
        block = torch.stack([columns[k][idx].to(dp.F64) for k in tens]).cpu().numpy()
        out = {k: block[i] for i, k in enumerate(tens)}
    for k, v in columns.items():
        if k not in out and v is not None:
            out[k] = np.zeros(n) if torch.is_tensor(v) else np.full(n, float(v))
    return out


def _semi_major_axis(mtot, P):
    return ((G * mtot * Msun) / (4 * pi ** 2) * (P * 86400) ** 2) ** (1 / 3)


_KEYS = ('M_s', 'R_s', 'u1', 'u2', 'P_orb', 'inc', 'b', 'R_p', 'ecc', 'argp', 'M_EB', 'R_EB',
         'fluxratio_EB', 'fluxratio_comp')


def _table(lb, dev, twin, M_host, R_host, u1, u2, P, mtot, incs, eccs, argps, cfr, rps=None,
           masses=None, radii=None, fluxratios=None):
    """This rank's half of a result dictionary (reference marginal_likelihoods.py:155-171): the
    rows of its best local draws, registered for the merge across ranks (_finished)."""
    idx = torch.as_tensor(np.asarray(lb.idx, dtype=np.int64), device=dev)
    n = len(lb.idx)
    got = _take_rows(dict(M_host=M_host, R_host=R_host, u1=u1, u2=u2, P=P, mtot=mtot, inc=incs,
                          ecc=eccs, argp=argps, cfr=cfr, rps=rps, masses=masses, radii=radii,
                          fluxratios=fluxratios), idx, dev)
    P_i = got["P"] * (2 if twin else 1)
    a_i = _semi_major_axis(got["mtot"], P_i)
    ecc, argp, inc, Rh = got["ecc"], got["argp"], got["inc"], got["R_host"]
    r = a_i * (1 - ecc ** 2) / (1 + ecc * np.sin(argp * np.pi / 180))
    zeros = np.zeros(n)
    local = {
        'M_s': got["M_host"], 'R_s': Rh, 'u1': got["u1"], 'u2': got["u2"], 'P_orb': P_i,
        'inc': inc, 'b': r * np.cos(inc * pi / 180) / (Rh * Rsun),
        'R_p': got.get("rps", zeros), 'ecc': ecc, 'argp': argp,
        'M_EB': got.get("masses", zeros), 'R_EB': got.get("radii", zeros),
        'fluxratio_EB': got.get("fluxratios", zeros),
        'fluxratio_comp': got["cfr"] if torch.is_tensor(cfr) else zeros,
    }
    return _dispatch.TableExchange(lb, local, _KEYS)


def _finished(exchange):
    from .marginal_likelihoods import ScenarioResult
    lnZ, n_pass, n_eval, merged = exchange.result()
    merged['lnZ'] = lnZ
    out = ScenarioResult(merged)
    out.n_pass, out.n_evaluated = n_pass, n_eval
    return out


def _run_tp(eng, dev, n, N, M_host, R_host, u1, u2, P, mtot, rps, incs, eccs, argps, cfr, lnprior,
            extra_mask, is_host):
    p = _dispatch._submit(eng, "tp_tensors", n,
                          dict(rp=rps, P_orb=P, inc=incs, ecc=eccs, argp=argps, mtot=mtot,
                               rhost=R_host, u1=u1, u2=u2, cfr=cfr, lnprior=lnprior),
                          extra_mask, is_host, N_SAMPLES)

    state = []

    def prepare():
        if not state:
            lb = _dispatch.gather_local(p.result(), N, eng)
            state.append(_table(lb, dev, False, M_host, R_host, u1, u2, P, mtot, incs, eccs,
                                argps, cfr, rps=rps))

    def table():
        prepare()
        return _finished(state[0])
    return _dispatch.deliver(table, prepare)


def _run_eb(eng, dev, n, N, M_host, R_host, u1, u2, P, mtot, incs, qs, eccs, argps, masses, radii,
            fluxratios, cfr, lnprior, extra_mask, is_host, scalar_loop=False):
    kw = {"scalar_loop": True} if scalar_loop else {}
    p = _dispatch._submit(eng, "eb_tensors", n,
                          dict(reb=radii, ebfr=fluxratios, q=qs, P_orb=P, inc=incs, ecc=eccs,
                               argp=argps, mtot=mtot, rhost=R_host, u1=u1, u2=u2, cfr=cfr,
                               lnprior=lnprior),
                          extra_mask, is_host, N_SAMPLES, **kw)
    common = (M_host, R_host, u1, u2, P, mtot, incs, eccs, argps, cfr)
    kw = dict(masses=masses, radii=radii, fluxratios=fluxratios)
    state, done = [], {}

    def prepare():   # both branches at the first request, in a fixed order
        if not state:
            r0, r1 = p.result()
            state.append(_table(_dispatch.gather_local(r0, N, eng), dev, False, *common, **kw))
            state.append(_table(_dispatch.gather_local(r1, N, eng), dev, True, *common, **kw))

    def table(b):
        prepare()
        if not done:   # (collectives inside when no CallGroup is open: fixed order)
            done[0], done[1] = _finished(state[0]), _finished(state[1])
        return done[b]
    return (_dispatch.deliver(lambda: table(0), prepare),
            _dispatch.deliver(lambda: table(1), prepare))


# ------------------------------------------------------------------------------- scenarios
def lnZ_TTP(time, flux, sigma, P_orb, M_s, R_s, Teff, Z, N=1000000, parallel=False,
            mission="TESS", flatpriors=False, exptime=0.00139, nsamples=20):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    u1, u2 = grid_for(mission).nearest(Z, Teff, _logg(M_s, R_s))
    rps, incs, eccs, argps = _draw_planet(n, M_s, flatpriors, P_mean, dev)
    return _run_tp(eng, dev, n, int(N), M_s, R_s, u1, u2, P, M_s, rps, incs, eccs, argps, 0.0,
                   None, None, False)


def lnZ_TEB(time, flux, sigma, P_orb, M_s, R_s, Teff, Z, N=1000000, parallel=False,
            mission="TESS", flatpriors=False, exptime=0.00139, nsamples=20):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    u1, u2 = grid_for(mission).nearest(Z, Teff, _logg(M_s, R_s))
    incs, qs, eccs, argps = _draw_binary(n, M_s, P_mean, dev)
    masses = qs * M_s
    radii, _ = dp.stellar_relations(masses, R_s, Teff)
    fluxratios = _fluxratio(masses, M_s)
    return _run_eb(eng, dev, n, int(N), M_s, R_s, u1, u2, P, M_s + masses, incs, qs, eccs, argps,
                   masses, radii, fluxratios, 0.0, None, None, False, scalar_loop=not parallel)


def lnZ_PTP(time, flux, sigma, P_orb, M_s, R_s, Teff, Z, plx, contrast_curve_file=None,
            filt="TESS", N=1000000, parallel=False, mission="TESS", flatpriors=False,
            exptime=0.00139, nsamples=20, molusc_file=None):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    u1, u2 = grid_for(mission).nearest(Z, Teff, _logg(M_s, R_s))
    qs_comp = _companion_q(n, M_s, molusc_file, dev)
    masses_comp = qs_comp * M_s
    cfr = _fluxratio(masses_comp, M_s)

    def cc_term():
        fr = _fluxratio(masses_comp, M_s, filt)
        return fr / (1 - fr)

    lnprior = _bound_prior(dp.lnprior_bound_TP, M_s, plx, n, dev, molusc_file,
                           contrast_curve_file, cfr / (1 - cfr), cc_term)
    rps, incs, eccs, argps = _draw_planet(n, M_s, flatpriors, P_mean, dev)
    return _run_tp(eng, dev, n, int(N), M_s, R_s, u1, u2, P, M_s, rps, incs, eccs, argps, cfr,
                   lnprior, qs_comp != 0.0, False)


def lnZ_PEB(time, flux, sigma, P_orb, M_s, R_s, Teff, Z, plx, contrast_curve_file=None,
            filt="TESS", N=1000000, parallel=False, mission="TESS", flatpriors=False,
            exptime=0.00139, nsamples=20, molusc_file=None):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    u1, u2 = grid_for(mission).nearest(Z, Teff, _logg(M_s, R_s))
    incs, qs, eccs, argps = _draw_binary(n, M_s, P_mean, dev)
    qs_comp = _companion_q(n, M_s, molusc_file, dev)
    masses = qs * M_s
    radii, _ = dp.stellar_relations(masses, R_s, Teff)
    fluxratios = _fluxratio(masses, M_s)
    masses_comp = qs_comp * M_s
    cfr = _fluxratio(masses_comp, M_s)

    def cc_term():
        fr = _fluxratio(masses_comp, M_s, filt)
        return fr / (1 - fr)

    lnprior = _bound_prior(dp.lnprior_bound_EB, M_s, plx, n, dev, molusc_file,
                           contrast_curve_file, cfr / (1 - cfr), cc_term)
    return _run_eb(eng, dev, n, int(N), M_s, R_s, u1, u2, P, M_s + masses, incs, qs, eccs, argps,
                   masses, radii, fluxratios, cfr, lnprior, qs_comp != 0.0, False, scalar_loop=not parallel)


def _companion_stars(n, M_s, R_s, Teff, Z, mission, qs_comp, Teff_cap):
    masses_comp = qs_comp * M_s
    radii_comp, Teffs_comp = dp.stellar_relations(masses_comp, R_s, Teff)
    loggs_comp = torch.log10(G * (masses_comp * Msun) / (radii_comp * Rsun) ** 2)
    cfr = _fluxratio(masses_comp, M_s)
    u1s, u2s = dp.ldc_at_Z_rounded(grid_for(mission), Z, Teffs_comp, loggs_comp, Teff_cap)
    return masses_comp, radii_comp, Teffs_comp, cfr, u1s, u2s


def lnZ_STP(time, flux, sigma, P_orb, M_s, R_s, Teff, Z, plx, contrast_curve_file=None,
            filt="TESS", N=1000000, parallel=False, mission="TESS", flatpriors=False,
            exptime=0.00139, nsamples=20, molusc_file=None):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    qs_comp = _companion_q(n, M_s, molusc_file, dev)
    masses_comp, radii_comp, _, cfr, u1s, u2s = _companion_stars(n, M_s, R_s, Teff, Z, mission,
                                                                 qs_comp, 10000)

    def cc_term():
        fr = _fluxratio(masses_comp, M_s, filt)
        return fr / (1 - fr)

    lnprior = _bound_prior(dp.lnprior_bound_TP, M_s, plx, n, dev, molusc_file,
                           contrast_curve_file, cfr / (1 - cfr), cc_term)
    rps, incs, eccs, argps = _draw_planet(n, masses_comp, flatpriors, P_mean, dev)
    return _run_tp(eng, dev, n, int(N), masses_comp, radii_comp, u1s, u2s, P, masses_comp, rps,
                   incs, eccs, argps, cfr, lnprior, qs_comp != 0.0, True)


def lnZ_SEB(time, flux, sigma, P_orb, M_s, R_s, Teff, Z, plx, contrast_curve_file=None,
            filt="TESS", N=1000000, parallel=False, mission="TESS", flatpriors=False,
            exptime=0.00139, nsamples=20, molusc_file=None):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    incs, qs, eccs, argps = _draw_binary(n, M_s, P_mean, dev)
    qs_comp = _companion_q(n, M_s, molusc_file, dev)
    masses_comp, radii_comp, Teffs_comp, cfr, u1s, u2s = _companion_stars(
        n, M_s, R_s, Teff, Z, mission, qs_comp, 13000)
    masses = qs * masses_comp
    radii, _ = dp.stellar_relations(masses, radii_comp, Teffs_comp)
    fluxratios = _fluxratio(masses, M_s)

    def cc_term():
        fr = _fluxratio(masses, M_s, filt)
        frc = _fluxratio(masses_comp, M_s, filt)
        return (frc / (1 - frc)) + (fr / (1 - fr))

    lnprior = _bound_prior(dp.lnprior_bound_EB, M_s, plx, n, dev, molusc_file,
                           contrast_curve_file,
                           (cfr / (1 - cfr)) + (fluxratios / (1 - fluxratios)), cc_term)
    return _run_eb(eng, dev, n, int(N), masses_comp, radii_comp, u1s, u2s, P,
                   masses_comp + masses, incs, qs, eccs, argps, masses, radii, fluxratios, cfr,
                   lnprior, qs_comp != 0.0, True, scalar_loop=not parallel)


def _randint(lo, hi, n, dev):
    return torch.randint(lo, hi, (n,), device=dev)


def lnZ_DTP(time, flux, sigma, P_orb, M_s, R_s, Teff, Z, Tmag, Jmag, Hmag, Kmag, trilegal_fname,
            contrast_curve_file=None, filt="TESS", N=1000000, parallel=False, mission="TESS",
            flatpriors=False, exptime=0.00139, nsamples=20):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    u1, u2 = grid_for(mission).nearest(Z, Teff, _logg(M_s, R_s))
    bg = _Background(trilegal_fname, Tmag, Jmag, Hmag, Kmag, dev)
    idxs = _randint(0, bg.N_comp - 1, n, dev)            # upper bound N_comp-1, as :1463
    cfr = bg.fluxratios[idxs]
    lnprior = _background_prior(bg, n, dev, contrast_curve_file,
                                2.5 * torch.log10(cfr / (1 - cfr)), bg.dmag(filt)[idxs])
    rps, incs, eccs, argps = _draw_planet(n, M_s, flatpriors, P_mean, dev)
    return _run_tp(eng, dev, n, int(N), M_s, R_s, u1, u2, P, M_s, rps, incs, eccs, argps, cfr,
                   lnprior, None, False)


def lnZ_DEB(time, flux, sigma, P_orb, M_s, R_s, Teff, Z, Tmag, Jmag, Hmag, Kmag, trilegal_fname,
            contrast_curve_file=None, filt="TESS", N=1000000, parallel=False, mission="TESS",
            flatpriors=False, exptime=0.00139, nsamples=20):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    u1, u2 = grid_for(mission).nearest(Z, Teff, _logg(M_s, R_s))
    incs, qs, eccs, argps = _draw_binary(n, M_s, P_mean, dev)
    masses = qs * M_s
    radii, _ = dp.stellar_relations(masses, R_s, Teff)
    fluxratios = _fluxratio(masses, M_s)
    bg = _Background(trilegal_fname, Tmag, Jmag, Hmag, Kmag, dev)
    idxs = _randint(0, bg.N_comp - 1, n, dev)
    cfr = bg.fluxratios[idxs]
    lnprior = _background_prior(bg, n, dev, contrast_curve_file,
                                2.5 * torch.log10(cfr / (1 - cfr)), bg.dmag(filt)[idxs])
    return _run_eb(eng, dev, n, int(N), M_s, R_s, u1, u2, P, M_s + masses, incs, qs, eccs, argps,
                   masses, radii, fluxratios, cfr, lnprior, None, False, scalar_loop=not parallel)


def lnZ_BTP(time, flux, sigma, P_orb, M_s, R_s, Teff, Tmag, Jmag, Hmag, Kmag, trilegal_fname,
            contrast_curve_file=None, filt="TESS", N=1000000, parallel=False, mission="TESS",
            flatpriors=False, exptime=0.00139, nsamples=20):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    bg = _Background(trilegal_fname, Tmag, Jmag, Hmag, Kmag, dev)
    radii_comp = bg.radii()
    u1c, u2c = bg.ldc(mission)
    idxs = _randint(0, bg.N_comp, n, dev)
    cfr = bg.fluxratios[idxs]
    lnprior = _background_prior(bg, n, dev, contrast_curve_file,
                                2.5 * torch.log10(cfr / (1 - cfr)), bg.dmag(filt)[idxs])
    host_masses = bg.masses[idxs]
    rps, incs, eccs, argps = _draw_planet(n, host_masses, flatpriors, P_mean, dev)
    extra = (bg.loggs[idxs] >= 3.5) & (bg.Teffs[idxs] <= 10000)
    return _run_tp(eng, dev, n, int(N), host_masses, radii_comp[idxs], u1c[idxs], u2c[idxs], P,
                   host_masses, rps, incs, eccs, argps, cfr, lnprior, extra, True)


def lnZ_BEB(time, flux, sigma, P_orb, M_s, R_s, Teff, Tmag, Jmag, Hmag, Kmag, trilegal_fname,
            contrast_curve_file=None, filt="TESS", N=1000000, parallel=False, mission="TESS",
            flatpriors=False, exptime=0.00139, nsamples=20):
    eng, dev, n = _begin(time, flux, sigma, exptime, nsamples, int(N))
    P, P_mean = _periods(P_orb, n, dev)
    incs, qs, eccs, argps = _draw_binary(n, M_s, P_mean, dev)
    bg = _Background(trilegal_fname, Tmag, Jmag, Hmag, Kmag, dev)
    radii_comp = bg.radii()
    u1c, u2c = bg.ldc(mission)
    idxs = _randint(0, bg.N_comp, n, dev)
    host_masses, host_radii = bg.masses[idxs], radii_comp[idxs]
    cfr = bg.fluxratios[idxs]
    masses = qs * host_masses
    radii, _ = dp.stellar_relations(masses, host_radii, bg.Teffs[idxs])

    def distance_corrected(band):
        cfr_band = bg.fluxratios_in(band)[idxs]
        bound = _fluxratio(host_masses, M_s, band)
        return _fluxratio(masses, M_s, band) * (cfr_band / bound), cfr_band

    fluxratios, _ = distance_corrected("TESS")
    if contrast_curve_file is None:
        dmag = 2.5 * torch.log10((cfr / (1 - cfr)) + (fluxratios / (1 - fluxratios)))
        lnprior = _background_prior(bg, n, dev, None, dmag, None)
    else:
        fr_cc, cfr_cc = distance_corrected(filt if filt in ("J", "H", "K") else "TESS")
        dmag = 2.5 * torch.log10((cfr_cc / (1 - cfr_cc)) + (fr_cc / (1 - fr_cc)))
        lnprior = _background_prior(bg, n, dev, contrast_curve_file, None, dmag)
    extra = (bg.loggs[idxs] >= 3.5) & (bg.Teffs[idxs] <= 10000)
    return _run_eb(eng, dev, n, int(N), host_masses, host_radii, u1c[idxs], u2c[idxs], P,
                   host_masses + masses, incs, qs, eccs, argps, masses, radii, fluxratios, cfr,
                   lnprior, extra, True, scalar_loop=not parallel)
