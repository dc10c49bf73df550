"""Prior draws and their transforms ON THE DEVICE (opt-in "device sampler" mode).

SURVEY.md section 8f row 1: after the light-curve path moved to the GPU, a `calc_probs` at
N = 1e6 spends ~2 s in host-side prior preparation (numpy's sequential Mersenne-Twister stream,
scipy's beta sampler, Python glue) and ~0.2 s on the GPU.  This module restates
`priors.py` / `funcs.py` / `_ldc.py` as float64 torch operations so that the draws are born in
HBM and are handed to `tri_eval_*_dev` by pointer; the host touches only scalars and the
100-row result tables.

It CANNOT be bit-identical to the reference: the deviates come from torch's Philox generator,
not from `np.random`, so results agree with the host mode only statistically (same
distributions, evidence equal within Monte-Carlo error; tests/test_device_sampler.py).  The
default mode therefore stays the host sampler, which reproduces the reference's arrays exactly.
The deterministic transforms are the same formulas as in `priors.py` / `funcs.py` (reference
priors.py:16-383, :580-1005; funcs.py:54-140, :222-238) and agree with them to rounding when
fed the same deviates.  One documented difference: contrast curves are interpolated by plain
bisection, while numpy.interp's result on a NON-monotonic curve depends on the order of the
queries (it carries a search guess from one element to the next), which no parallel evaluation
can reproduce.

torch is used here as the device-side array library for preparation work (<1 % of the GPU time);
the hot path stays in csrc/.
"""
import math

import numpy as np
import torch

from . import funcs
from ._constants import G, Msun, au, pi

F64 = torch.float64


def _t(x, dev):
    return torch.as_tensor(x, dtype=F64, device=dev)


# ------------------------------------------------------------------------------- samplers
def rand(N, dev):
    return torch.rand(N, dtype=F64, device=dev)


def _piecewise_powerlaw(x, select, edges, powers, amps):
    """Inverse-CDF transform onto a broken power law (priors._piecewise_powerlaw)."""
    nseg = len(powers)
    integrals = []
    for k in range(nseg):
        p1 = powers[k] + 1
        integrals.append(amps[k] * (edges[k + 1] ** p1 - edges[k] ** p1) / p1)
    cum = list(np.cumsum(integrals))
    norm = 1 / cum[-1]
    out = x.clone()
    for k in range(nseg):
        m = x <= norm * cum[k]
        if k > 0:
            m = m & (x > norm * cum[k - 1])
        if select is not None:
            m = m & select
        p1 = powers[k] + 1
        u = x / norm
        for j in range(k):
            u = u - integrals[j]
        val = (torch.clamp(u * p1 / amps[k] + edges[k] ** p1, min=1e-300)) ** (1 / p1)
        out = torch.where(m, val, out)
    return out


def sample_rp(x, M_host, flatpriors):
    if flatpriors:
        return x / (1 / 19.5) + 0.5
    edges = (0.5, 3.0, 6.0, 20.0)
    M_host = torch.as_tensor(M_host, dtype=F64, device=x.device)
    out = x
    for powers, select in (((0.0, -4.0, -0.5), M_host > 0.45), ((0.0, -7.0, -0.5), M_host <= 0.45)):
        p1, p2, p3 = powers
        A1 = edges[1] ** p1 / edges[1] ** p2
        A2 = edges[2] ** p2 / edges[2] ** p3
        sel = select.expand_as(x) if select.ndim == 0 else select
        out = torch.where(sel, _piecewise_powerlaw(x, sel, edges, powers, (1.0, A1, A2 * A1)), out)
    return out


def sample_inc(x):
    return torch.acos(1.0 - x) * (180 / math.pi)


def sample_ecc(N, planet, P_mean, dev):
    if planet:
        a = torch._standard_gamma(torch.full((N,), 0.867, dtype=F64, device=dev))
        b = torch._standard_gamma(torch.full((N,), 3.030, dtype=F64, device=dev))
        return a / (a + b)
    expo = 0.2 if P_mean <= 10 else 0.6
    return rand(N, dev) ** (1 / expo)


def sample_w(x):
    return x * 360


def _sample_mass_ratio(x, M_s, p2, F_twin):
    p1 = 0.3
    e2 = p2 + 1
    if M_s >= 0.3:
        q_lo = 0.1 if M_s >= 1.0 else 0.1 / M_s
        A1 = (0.3 ** p1) / (0.3 ** p2)
        A2 = (1 + F_twin / (1 - F_twin) * ((1.0 ** e2 - 0.3 ** e2) / e2)
              / ((1.0 ** e2 - 0.95 ** e2) / e2))
        return _piecewise_powerlaw(x, None, (q_lo, 0.3, 0.95, 1.0), (p1, p2, p2),
                                   (1.0, A1, A2 * A1))
    if M_s > 0.1:
        q_lo = 0.1 / M_s
        A2 = (1 + F_twin / (1 - F_twin) * ((1.0 ** e2 - q_lo ** e2) / e2)
              / ((1.0 ** e2 - 0.95 ** e2) / e2))
        return _piecewise_powerlaw(x, None, (q_lo, 0.95, 1.0), (p2, p2), (1.0, A2))
    return torch.ones_like(x)


def sample_q(x, M_s):
    return _sample_mass_ratio(x, M_s, -0.5, 0.30)


def sample_q_companion(x, M_s):
    return _sample_mass_ratio(x, M_s, -0.95, 0.05)


# ------------------------------------------------------------------------------- relations
_spline_cache = {}


def _spline_tensors(spl, dev):
    key = (id(spl), str(dev))
    if key not in _spline_cache:
        t, c, k = spl._eval_args
        _spline_cache[key] = (_t(t, dev), _t(c, dev), int(k))
    return _spline_cache[key]


def splev(spl, x):
    """FITPACK B-spline (t, c, k) at x with extrapolation (de Boor's recurrence).  CUDA tensors
    go through the library's splev kernel (one launch); CPU tensors, which only the tests use,
    through the same recurrence spelled in torch ops."""
    t, c, k = _spline_tensors(spl, x.device)
    n = t.numel()
    if x.is_cuda:
        from . import _cabi
        xc = x.contiguous()
        y = torch.empty_like(xc)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _cabi.check(_cabi.load().tri_dev_splev(t.data_ptr(), c.data_ptr(), n, k, xc.data_ptr(),
                                               y.data_ptr(), xc.numel(), stream))
        return y
    l = torch.searchsorted(t, x, right=True) - 1          # t[l] <= x < t[l+1]
    l = torch.clamp(l, k, n - k - 2)
    h = [torch.ones_like(x)] + [torch.zeros_like(x) for _ in range(k)]
    for j in range(1, k + 1):
        hh = [v.clone() for v in h[:j]]
        h[0] = torch.zeros_like(x)
        for i in range(j):
            ti = t[l + i + 1]
            tj = t[l + i + 1 - j]
            f = hh[i] / (ti - tj)
            h[i] = h[i] + f * (ti - x)
            h[i + 1] = f * (x - tj)
    sp = torch.zeros_like(x)
    for j in range(k + 1):
        sp = sp + c[l - k + j] * h[j]
    return sp


def stellar_relations(Masses, max_Radii, max_Teffs):
    hot = Masses > 0.63
    R = torch.where(hot, splev(funcs._hot_R, Masses), splev(funcs._cool_R, Masses))
    T = torch.where(hot, splev(funcs._hot_T, Masses), splev(funcs._cool_T, Masses))
    R = torch.minimum(R, torch.as_tensor(max_Radii, dtype=F64, device=Masses.device))
    T = torch.minimum(T, torch.as_tensor(max_Teffs, dtype=F64, device=Masses.device))
    return torch.clamp(R, min=0.1), torch.clamp(T, min=2800.0)


def flux_relation(Masses, filt="TESS"):
    return 10 ** splev(funcs._FLUX_SPLINES[filt], Masses)


def flux_relation_scalar(M, filt="TESS"):
    return float(funcs.flux_relation(np.array([M]), filt)[0])


def interp(x, xp, fp):
    """Piecewise-linear interpolation with end clamping (numpy.interp for monotonic xp; plain
    bisection otherwise, see the module docstring)."""
    n = xp.numel()
    if n == 1:
        return fp[0].expand_as(x).clone()
    j = torch.clamp(torch.searchsorted(xp, x, right=True) - 1, 0, n - 2)
    slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j])
    y = slope * (x - xp[j]) + fp[j]
    y = torch.where(x >= xp[-1], fp[-1], y)
    return torch.where(x <= xp[0], fp[0], y)


def _bound_companion_lnprior(M_s, plx, delta_mags, separations, contrasts, first_decade):
    if np.isnan(plx):
        plx = 0.1
    d = 1000 / plx
    seps = d * interp(delta_mags, contrasts, separations)
    M_act = M_s
    if not (M_s >= 1.0):
        M_s = 1.0
    lm = math.log10(M_s)
    f1 = 0.020 + 0.04 * lm + 0.07 * lm ** 2
    f2 = 0.039 + 0.07 * lm + 0.01 * lm ** 2
    f3 = 0.078 - 0.05 * lm + 0.04 * lm ** 2
    alpha, dlogP = 0.018, 0.7
    slope = f2 - f1 - alpha * dlogP
    slope2 = f3 - f2 - alpha * dlogP
    t2 = 0.5 * (2.0 * f1 + slope)
    t3 = 0.5 * alpha * (3.4 ** 2 - 5.4 * 3.4 + 6.8) + f2 * (3.4 - 2.0)
    t4 = (alpha * dlogP * (5.5 - 3.4) + f2 * (5.5 - 3.4)
          + slope2 * (0.238095 * 5.5 ** 2 - 0.952381 * 5.5 + 0.485714))
    t5 = f3 * (3.33333 - 17.3566 * math.exp(-0.3 * 8.0))
    max_Porbs = torch.sqrt((4 * pi ** 2) / (G * M_s * Msun) * (seps * au) ** 3) / 86400
    lp = torch.log10(max_Porbs)
    t2p = 0.5 * (lp - 1.0) * (2.0 * f1 + slope * (lp - 1.0))
    t3p = 0.5 * alpha * (lp ** 2 - 5.4 * lp + 6.8) + f2 * (lp - 2.0)
    t4p = (alpha * dlogP * (lp - 3.4) + f2 * (lp - 3.4)
           + slope2 * (0.238095 * lp ** 2 - 0.952381 * lp + 0.485714))
    t5p = f3 * (3.33333 - 17.3566 * torch.exp(-0.3 * lp))
    z = torch.zeros_like(lp)
    if first_decade:
        f = torch.where(lp >= 8.0, z + (t2 + t3 + t4 + t5),
            torch.where(lp >= 5.5, t2 + t3 + t4 + t5p,
            torch.where(lp >= 3.4, t2 + t3 + t4p,
            torch.where(lp >= 2.0, t2 + t3p,
            torch.where(lp >= 1.0, t2p, z)))))
    else:
        f = torch.where(lp >= 8.0, z + (t4 + t5),
            torch.where(lp >= 5.5, t4 + t5p,
            torch.where(lp >= 3.4, t4p, z)))
    if M_act >= 1.0:
        return torch.log(f)
    f_act = torch.clamp(0.65 * f + 0.35 * f * M_act, min=0.0)
    return torch.log(f_act)


def lnprior_bound_TP(M_s, plx, delta_mags, separations, contrasts):
    return _bound_companion_lnprior(M_s, plx, delta_mags, separations, contrasts, False)


def lnprior_bound_EB(M_s, plx, delta_mags, separations, contrasts):
    return _bound_companion_lnprior(M_s, plx, delta_mags, separations, contrasts, True)


def lnprior_background(N_comp, delta_mags, separations, contrasts):
    seps = interp(delta_mags, contrasts, separations)
    return torch.log((N_comp / 0.1) * (1 / 3600) ** 2 * seps ** 2)


def clip_prior(lnprior, delta_mags):
    lnprior = torch.clamp(lnprior, max=0.0)
    return torch.where(delta_mags > 0.0, torch.full_like(lnprior, -math.inf), lnprior)


# ------------------------------------------------------------------------------- LDC grids
_ldc_cache = {}


def ldc_at_Z_rounded(grid, Z, Teffs, loggs, Teff_cap):
    """_ldc.LdcGrid.at_Z_rounded on the device: dense (Teff, logg) table at the target's Z."""
    dev = Teffs.device
    key = (id(grid), float(Z), str(dev))
    if key not in _ldc_cache:
        at_Z = grid.Zs == grid.Zs[np.abs(grid.Zs - Z).argmin()]
        T_at, g_at = grid.Teffs[at_Z], grid.loggs[at_Z]
        tab1 = np.full((27, 4), np.nan)
        tab2 = np.full((27, 4), np.nan)
        it = ((T_at - 3500) // 250).astype(int)
        ig = np.round((g_at - 3.5) / 0.5).astype(int)
        tab1[it, ig] = grid.u1s[at_Z]
        tab2[it, ig] = grid.u2s[at_Z]
        _ldc_cache[key] = (_t(tab1, dev), _t(tab2, dev))
    tab1, tab2 = _ldc_cache[key]
    rg = torch.clamp(torch.round(loggs / 0.5) * 0.5, 3.5, 5.0)
    rT = torch.clamp(torch.round(Teffs / 250) * 250, 3500, Teff_cap)
    it = ((rT - 3500) / 250).long()
    ig = torch.round((rg - 3.5) / 0.5).long()
    if Teff_cap > 10000 and int(it.max()) > 26:   # (a device read-back: only where it can fire)
        # the reference's `.item()` raises for nodes beyond the grid (Teff clamp 13000, :1181)
        raise ValueError("can only convert an array of size 1 to a Python scalar")
    return tab1[it, ig], tab2[it, ig]
