"""ORACLE (test infrastructure, not product code): CPU port of the scenario seam.

Same call surface as triceratops_b200.engine.Engine, computed the way the reference computes it:
the geometry / mask expressions of marginal_likelihoods.py:107-123 (TP-type) and :254-299
(EB-type) in numpy, the per-draw likelihoods through oracle/coracle.py (C restatement of
likelihoods.py:302-587, OpenMP over draws), and _log_mean_exp (_numerics.py:12-51).
PARITY UNPINNED with respect to real pytransit (see oracle/quadmodel.py).

Users: tests/ (as the checker, and as a stand-in that lets CPU tests exercise the Python host
layer and the multi-rank combine over gloo), __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the triceratops_b200 package.
"""
import math

import numpy as np

from oracle import coracle

# astropy >= 4.0 `constants.X.cgs.value` (CODATA 2018 / IAU 2015 nominal), which the reference
# reads at likelihoods.py:17-21 and marginal_likelihoods.py:13-17 -- the checker's own copy
G = 6.6743e-08
Msun = 1.988409870698051e+33
Rsun = 69570000000.0
Rearth = 637810000.0
pi = np.pi
ln2pi = np.log(2 * pi)


class _Res:
    pass


def _full(x, N):
    return np.full(N, float(x)) if np.ndim(x) == 0 else np.asarray(x, dtype=float)


def _lse_record(lnw, N, res):
    fin = np.isfinite(lnw)
    res.n_posinf = int(np.isposinf(lnw).sum())
    res.n_finite = int(fin.sum())
    if res.n_finite:
        res.m = float(lnw[fin].max())
        res.s = float(np.exp(lnw[fin] - res.m).sum())
    else:
        res.m, res.s = -math.inf, 0.0
    if res.n_posinf:
        res.lnZ = math.inf
    elif not res.n_finite:
        res.lnZ = -math.inf
    else:
        res.lnZ = res.m + math.log(res.s) - math.log(N)
    return res


def tp_mask(N, rp, P_orb, inc, ecc, argp, mtot, rhost, extra_mask=None):
    """Geometric mask of a TP-type scenario (marginal_likelihoods.py:107-123) and the
    semi-major axis [cm], numpy over all N draws."""
    rp, P, inc, ecc, argp, mtot, rhost = [
        _full(x, N) for x in (rp, P_orb, inc, ecc, argp, mtot, rhost)]
    a = ((G * mtot * Msun) / (4 * pi ** 2) * (P * 86400) ** 2) ** (1 / 3)
    e_corr = (1 + ecc * np.sin(argp * pi / 180)) / (1 - ecc ** 2)
    Ptra = (rp * Rearth + rhost * Rsun) / a * e_corr
    coll = (rp * Rearth + rhost * Rsun) > a * (1 - ecc)
    inc_min = np.full(N, 90.)
    ok = Ptra <= 1.
    inc_min[ok] = np.arccos(Ptra[ok]) * 180. / pi
    mask = (inc >= inc_min) & (coll == False)  # noqa: E712
    if extra_mask is not None:
        mask &= np.asarray(extra_mask, bool)
    return mask, a


def eb_masks(N, reb, q, P_orb, inc, ecc, argp, mtot, rhost, extra_mask=None, scalar_loop=False):
    """Masks and semi-major axes of the EB (q < 0.95, period P) and EBx2P (q >= 0.95, period
    2P) branches (marginal_likelihoods.py:254-299): ((mask, a), (mask_twin, a_twin)).
    scalar_loop: the reference's parallel=False loop (:313-339) `continue`s past BOTH branches
    of a draw whose period-P transit probability exceeds 1."""
    reb, q, P, inc, ecc, argp, mtot, rhost = [
        _full(x, N) for x in (reb, q, P_orb, inc, ecc, argp, mtot, rhost)]
    e_corr = (1 + ecc * np.sin(argp * pi / 180)) / (1 - ecc ** 2)
    out = []
    for twin in (False, True):
        Pk = 2 * P if twin else P
        a = ((G * mtot * Msun) / (4 * pi ** 2) * (Pk * 86400) ** 2) ** (1 / 3)
        Ptra = (reb * Rsun + rhost * Rsun) / a * e_corr
        if twin:
            coll = (2 * rhost * Rsun) > a * (1 - ecc)
        else:
            coll = (reb * Rsun + rhost * Rsun) > a * (1 - ecc)
        inc_min = np.full(N, 90.)
        ok = Ptra <= 1.
        inc_min[ok] = np.arccos(Ptra[ok]) * 180. / pi
        mask = (inc >= inc_min) & (coll == False) & ((q >= 0.95) if twin else (q < 0.95))  # noqa: E712
        if extra_mask is not None:
            mask &= np.asarray(extra_mask, bool)
        if twin and scalar_loop:
            mask &= Ptra_P <= 1.
        if not twin:
            Ptra_P = Ptra
        out.append((mask, a))
    return tuple(out)


class OracleEngine:
    device = -1

    @property
    def torch_device(self):
        """Device-sampler mode on the CPU stand-in: the prior draws are torch CPU tensors."""
        import torch
        return torch.device("cpu")

    def set_lightcurve(self, time, flux, sigma, exptime, nsamples):
        self.lc = (np.asarray(time, float), np.asarray(flux, float), float(sigma),
                   float(exptime), int(nsamples))

    def eval_tp(self, N, rp, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr, lnprior=None,
                extra_mask=None, companion_is_host=False, want_lnL=True, want_mask=False,
                n_best=0):
        t, f, s, exptime, ns = self.lc
        rp, P, inc, ecc, argp, mtot, rhost, u1, u2, cfr = [
            _full(x, N) for x in (rp, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr)]
        mask, a = tp_mask(N, rp, P, inc, ecc, argp, mtot, rhost, extra_mask)
        lnL = np.full(N, -np.inf)
        if mask.any():
            lnL[mask] = -0.5 * ln2pi - np.log(s) - coracle.lnL_TP_p(
                t, f, s, rp[mask], P[mask], inc[mask], a[mask], rhost[mask], u1[mask], u2[mask],
                ecc[mask], argp[mask], cfr[mask], companion_is_host, exptime, ns)
        res = _Res()
        res.N, res.lnL, res.mask, res.n_pass, res.n_stamps = N, lnL, mask, int(mask.sum()), 0
        return _lse_record(lnL if lnprior is None else lnL + _full(lnprior, N), N, res)

    def eval_eb(self, N, reb, ebfr, q, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr,
                lnprior=None, extra_mask=None, companion_is_host=False, want_lnL=True,
                want_mask=False, n_best=0, scalar_loop=False):
        t, f, s, exptime, ns = self.lc
        reb, ebfr, q, P, inc, ecc, argp, mtot, rhost, u1, u2, cfr = [
            _full(x, N) for x in (reb, ebfr, q, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr)]
        out = []
        masks = eb_masks(N, reb, q, P, inc, ecc, argp, mtot, rhost, extra_mask, scalar_loop)
        for twin in (False, True):
            Pk = 2 * P if twin else P
            mask, a = masks[int(twin)]
            lnL = np.full(N, -np.inf)
            if mask.any():
                fn = coracle.lnL_EB_twin_p if twin else coracle.lnL_EB_p
                lnL[mask] = -0.5 * ln2pi - np.log(s) - fn(
                    t, f, s, reb[mask], ebfr[mask], Pk[mask], inc[mask], a[mask], rhost[mask],
                    u1[mask], u2[mask], ecc[mask], argp[mask], cfr[mask], companion_is_host,
                    exptime, ns, scalar_rule=scalar_loop)
            res = _Res()
            res.N, res.lnL, res.mask, res.n_pass, res.n_stamps = N, lnL, mask, int(mask.sum()), 0
            out.append(_lse_record(lnL if lnprior is None else lnL + _full(lnprior, N), N, res))
        return tuple(out)

    # ---- device-sampler entry points: torch (CPU) tensors in, numpy evaluation ----
    @staticmethod
    def _np(x):
        try:
            import torch
            if torch.is_tensor(x):
                return x.detach().cpu().numpy()
        except Exception:
            pass
        return x

    def _with_top(self, res, n_best):
        fin = np.flatnonzero(np.isfinite(res.lnL))
        order = fin[np.lexsort((fin, -res.lnL[fin]))][:n_best]
        res.top_idx, res.top_lnL = order, res.lnL[order]
        res.n_evaluated, res.branch = int(fin.size), 0
        return res

    def eval_tp_tensors(self, N, cols, extra_mask=None, companion_is_host=False, n_best=100):
        c = {k: self._np(v) for k, v in cols.items()}
        res = self.eval_tp(N, extra_mask=self._np(extra_mask),
                           companion_is_host=companion_is_host, **c)
        return self._with_top(res, n_best)

    def eval_eb_tensors(self, N, cols, extra_mask=None, companion_is_host=False, n_best=100,
                        scalar_loop=False):
        c = {k: self._np(v) for k, v in cols.items()}
        r0, r1 = self.eval_eb(N, extra_mask=self._np(extra_mask),
                              companion_is_host=companion_is_host, scalar_loop=scalar_loop, **c)
        return self._with_top(r0, n_best), self._with_top(r1, n_best)

    # ---- simulate seam (likelihoods.py:302-439 and the scalar :27-160), via the C model ----
    def _rows(self, npts, k, P, a_cm, R_s, inc, ecc, w_deg, u1, u2, exptime, ns):
        t = self.lc[0]
        out = np.empty((len(k), npts))
        for i in range(len(k)):
            out[i] = coracle.model(t, k[i], P[i], a_cm[i] / (R_s[i] * Rsun), inc[i] * (pi / 180.),
                                   ecc[i], w_deg[i] * (pi / 180.), u1[i], u2[i], exptime, ns)
        return out

    def simulate_tp(self, npts, R_p, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr,
                    companion_is_host):
        _, _, _, exptime, ns = self.lc
        n = np.size(R_p)
        R_p, P, inc, a, R_s, u1, u2, ecc, argp, cfr = [
            _full(x, n) for x in (R_p, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr)]
        flux = self._rows(npts, R_p * Rearth / (R_s * Rsun), P, a, R_s, inc, ecc, 90 - argp, u1,
                          u2, exptime, ns)
        F_comp = (cfr / (1 - cfr)).reshape(-1, 1)
        D = 1 / F_comp if companion_is_host else F_comp / 1
        return (flux + D) / (1 + D)

    def simulate_eb(self, npts, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr,
                    companion_is_host, scalar_rule=False):
        _, _, _, exptime, ns = self.lc
        n = np.size(R_EB)
        R_EB, fr, P, inc, a, R_s, u1, u2, ecc, argp, cfr = [
            _full(x, n) for x in (R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr)]
        k = R_EB / R_s
        if scalar_rule:
            k[np.abs(k - 1.0) < 1e-6] *= 0.999
            ks = 1 / k
        else:
            k[(k - 1.0) < 1e-6] *= 0.999
            ks = R_s / R_EB
            ks[(ks - 1.0) < 1e-6] *= 0.999
        flux = self._rows(npts, k, P, a, R_s, inc, ecc, 90 - argp, u1, u2, exptime, ns)
        tsec = np.linspace(-0.05, 0.05, 25)
        sec = np.empty(n)
        for i in range(n):
            sec[i] = coracle.model(tsec, ks[i], P[i], a[i] / (R_s[i] * Rsun), inc[i] * (pi / 180.),
                                   ecc[i], (90 - argp[i] + 180) * (pi / 180.), u1[i], u2[i],
                                   0.0, 1).min()
        F_comp = (cfr / (1 - cfr)).reshape(-1, 1)
        F_EB = (fr / (1 - fr)).reshape(-1, 1)
        sec = sec.reshape(-1, 1)
        if companion_is_host:
            flux = (flux + F_EB / F_comp) / (1 + F_EB / F_comp)
            sec = (sec + F_comp / F_EB) / (1 + F_comp / F_EB)
            D = 1 / (F_comp + F_EB)
        else:
            flux = (flux + F_EB / 1) / (1 + F_EB / 1)
            sec = (sec + 1 / F_EB) / (1 + 1 / F_EB)
            D = F_comp / (1 + F_EB)
        flux = (flux + D) / (1 + D)
        return flux, (1 - (sec + D) / (1 + D))[:, 0]

    def lnl_tp(self, R_p, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr, companion_is_host):
        t, f, s, exptime, ns = self.lc
        return coracle.lnL_TP_p(t, f, s, R_p, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr,
                                companion_is_host, exptime, ns)

    def lnl_eb(self, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr,
               companion_is_host, twin):
        t, f, s, exptime, ns = self.lc
        fn = coracle.lnL_EB_twin_p if twin else coracle.lnL_EB_p
        return fn(t, f, s, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp, cfr,
                  companion_is_host, exptime, ns)


_instance = OracleEngine()


_saved_get_engine = None


def install(instance=None):
    """Route triceratops_b200's host layer to the oracle stand-in by patching
    `_dispatch.get_engine` from outside (tests, bench.py's CPU arm)."""
    global _saved_get_engine
    from triceratops_b200 import _dispatch
    if _saved_get_engine is None:
        _saved_get_engine = _dispatch.get_engine
    eng = instance if instance is not None else _instance
    _dispatch.get_engine = lambda: eng
    return eng


def uninstall():
    global _saved_get_engine
    from triceratops_b200 import _dispatch
    if _saved_get_engine is not None:
        _dispatch.get_engine = _saved_get_engine
        _saved_get_engine = None
