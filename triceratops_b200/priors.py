"""Prior samplers and companion-rate priors: the host side of the marginal-likelihood path.

Mirrors triceratops/priors.py of the reference (samplers :16-383, companion priors :580-1005).
The draws are made on the host from numpy's global RNG in the reference's call order, so a given
`np.random.seed` yields the same arrays as the reference, bit for bit; the GPU consumes them
unmasked.  The unused `lnprior_Mstar_*` / `lnprior_Porb_*` functions of the reference
(priors.py:386-577, dead code) are not reproduced.

Implementation note: the reference writes each broken power law out longhand; here one
inverse-CDF helper serves all of them.  The per-element floating-point operations (and their
order) are the same as in the reference expressions, which is what keeps the arrays identical.
"""
import numpy as np
from scipy.stats import beta, powerlaw

from . import _hostpar
from ._constants import G, Msun, au, pi
from .funcs import separation_at_contrast


def powerlaw_tables(edges, powers, amps):
    """(segment integrals, their running sums, 1 / total) of a broken power law -- the scalars
    of the inverse-CDF transform, shared by the numpy path below and csrc/host_blocks.c."""
    nseg = len(powers)
    integrals = []
    for k in range(nseg):
        p1 = powers[k] + 1
        integrals.append(amps[k] * (edges[k + 1] ** p1 - edges[k] ** p1) / p1)
    cum = []
    tot = integrals[0]
    cum.append(tot)
    for k in range(1, nseg):
        tot = tot + integrals[k]
        cum.append(tot)
    return integrals, cum, 1 / tot


def _piecewise_powerlaw(x, select, edges, powers, amps):
    """In-place inverse-CDF transform of uniform deviates onto a broken power law.

    Segment k spans [edges[k], edges[k+1]] with density amps[k] * r**powers[k].  Only elements
    where `select` is true are touched; deviates beyond the last CDF knot are left as they are
    (as in the reference, priors.py:54-111).
    """
    nseg = len(powers)
    integrals, cum, norm = powerlaw_tables(edges, powers, amps)

    def transform(xv, sel):
        # all segment masks are taken from the untouched deviates before any is overwritten
        masks = []
        for k in range(nseg):
            m = xv <= norm * cum[k]
            if k > 0:
                m = (xv > norm * cum[k - 1]) & m
            if sel is not None:
                m = m & sel
            masks.append(m)
        for k in range(nseg):
            m = masks[k]
            p1 = powers[k] + 1
            u = xv[m] / norm
            for j in range(k):
                u = u - integrals[j]
            xv[m] = (u * p1 / amps[k] + edges[k] ** p1) ** (1 / p1)

    # element-wise and in place: chunks of the same array can be transformed concurrently
    _hostpar.pmap(transform, x.shape[0], x, select)
    return x


def sample_rp(x, M_s, flatpriors):
    """Planet radii [R_earth] from uniform deviates x, conditioned on host mass (priors.py:16-116)."""
    if flatpriors == False:  # noqa: E712  (the reference accepts numpy bools here)
        # two mass regimes with different middle slopes; their selections are disjoint, so the
        # second pass still sees untouched deviates
        (hi, lo) = rp_specs()
        _piecewise_powerlaw(x, M_s > 0.45, *hi)
        _piecewise_powerlaw(x, M_s <= 0.45, *lo)
        return x
    elif flatpriors == True:  # noqa: E712
        return x / RP_FLAT_A + 0.5


RP_FLAT_A = 1 / 19.5


def rp_specs():
    """(edges, powers, amps) of the planet-radius law for hosts above / not above 0.45 Msun."""
    edges = (0.5, 3.0, 6.0, 20.0)
    out = []
    for powers in ((0.0, -4.0, -0.5), (0.0, -7.0, -0.5)):
        p1, p2, p3 = powers
        A1 = edges[1] ** p1 / edges[1] ** p2
        A2 = edges[2] ** p2 / edges[2] ** p3
        out.append((edges, powers, (1.0, A1, A2 * A1)))
    return tuple(out)


def sample_inc(x, lower=0, upper=90):
    """Inclinations [deg], isotropic between lower and upper (priors.py:119-132)."""
    c_lo = np.cos(lower * np.pi / 180)
    norm = 1 / (c_lo - np.cos(upper * np.pi / 180))
    return _hostpar.pmap_concat(lambda xv: np.arccos(c_lo - xv / norm) * 180 / np.pi,
                                x.shape[0], x)


def sample_ecc(x, planet, P_orb):
    """Eccentricities: Beta(0.867, 3.03) for planets, power law for binaries (priors.py:134-155).

    As in the reference the deviates x are ignored except for their length; the draw comes from
    scipy.stats on numpy's global RNG."""
    size = len(x)
    if planet == True:  # noqa: E712
        return beta.rvs(0.867, 3.030, size=size)
    return powerlaw.rvs(0.2 if P_orb <= 10 else 0.6, size=size)


def sample_w(x):
    """Arguments of periastron [deg] (priors.py:157-166)."""
    return x * 360


Q_LAW = (-0.5, 0.30)             # (p2, F_twin) of sample_q
Q_COMPANION_LAW = (-0.95, 0.05)   # ... and of sample_q_companion


def mass_ratio_spec(M_s, p2, F_twin):
    """(edges, powers, amps) of the mass-ratio law for primary mass M_s, or None when every
    mass ratio is 1 (M_s <= 0.1 Msun)."""
    p1 = 0.3
    e2 = p2 + 1
    if M_s >= 1.0 or (M_s < 1.0) & (M_s >= 0.3):
        q_lo = 0.1 if M_s >= 1.0 else 0.1 / M_s
        A1 = (0.3 ** p1) / (0.3 ** p2)
        A2 = (1 + (F_twin) / (1 - F_twin)
              * ((1.0 ** e2 - 0.3 ** e2) / e2)
              / ((1.0 ** e2 - 0.95 ** e2) / e2))
        return (q_lo, 0.3, 0.95, 1.0), (p1, p2, p2), (1.0, A1, A2 * A1)
    if (M_s < 0.3) & (M_s > 0.1):
        q_lo = 0.1 / M_s
        A2 = (1 + (F_twin) / (1 - F_twin)
              * ((1.0 ** e2 - q_lo ** e2) / e2)
              / ((1.0 ** e2 - 0.95 ** e2) / e2))
        return (q_lo, 0.95, 1.0), (p2, p2), (1.0, A2)
    return None


def _sample_mass_ratio(x, M_s, p2, F_twin):
    """Shared body of sample_q / sample_q_companion (priors.py:168-274 / :277-383)."""
    spec = mass_ratio_spec(M_s, p2, F_twin)
    if spec is None:
        return np.full(len(x), 1.0)
    return _piecewise_powerlaw(x, None, *spec)


def sample_q(x, M_s):
    """Mass ratios of short-period binaries (priors.py:168-274)."""
    return _sample_mass_ratio(x, M_s, *Q_LAW)


def sample_q_companion(x, M_s):
    """Mass ratios of long-period bound companions (priors.py:277-383)."""
    return _sample_mass_ratio(x, M_s, *Q_COMPANION_LAW)


def bound_constants(M_s, plx):
    """The scalars of the bound-companion prior (period-distribution coefficients of Moe &
    Di Stefano 2017 for primary mass M_s and the segment integrals), shared by the numpy path
    below and csrc/host_blocks.c."""
    if np.isnan(plx):
        plx = 0.1
    d = 1000 / plx
    M_act = M_s
    if not (M_s >= 1.0):
        M_s = 1.0
    lm = np.log10(M_s)
    f1 = 0.020 + 0.04 * lm + 0.07 * (lm) ** 2
    f2 = 0.039 + 0.07 * lm + 0.01 * (lm) ** 2
    f3 = 0.078 - 0.05 * lm + 0.04 * (lm) ** 2
    alpha = 0.018
    dlogP = 0.7
    slope = f2 - f1 - alpha * dlogP
    slope2 = f3 - f2 - alpha * dlogP
    t2 = 0.5 * (2.0 - 1.0) * (2.0 * f1 + slope * (2.0 - 1.0))
    t3 = 0.5 * alpha * (3.4 ** 2 - 5.4 * 3.4 + 6.8) + f2 * (3.4 - 2.0)
    t4 = (alpha * dlogP * (5.5 - 3.4) + f2 * (5.5 - 3.4)
          + slope2 * (0.238095 * 5.5 ** 2 - 0.952381 * 5.5 + 0.485714))
    t5 = f3 * (3.33333 - 17.3566 * np.exp(-0.3 * 8.0))
    return dict(d=d, M_act=M_act, M_s=M_s, f1=f1, f2=f2, f3=f3, alpha=alpha, dlogP=dlogP,
                slope=slope, slope2=slope2, t2=t2, t3=t3, t4=t4, t5=t5)


def _bound_companion_lnprior(M_s, plx, delta_mags, separations, contrasts, first_decade):
    """ln of the fraction of targets with a bound companion inside the contrast-curve limit.

    Piecewise period distribution of Moe & Di Stefano (2017) integrated from log P = 1 (EB
    scenarios, first_decade=True; priors.py:784-984) or from log P = 3.4 (planet scenarios,
    first_decade=False; priors.py:580-782) up to the period of the widest allowed orbit.
    """
    K = bound_constants(M_s, plx)
    d, M_act, M_s = K["d"], K["M_act"], K["M_s"]
    f1, f2, f3, alpha, dlogP = K["f1"], K["f2"], K["f3"], K["alpha"], K["dlogP"]
    slope, slope2 = K["slope"], K["slope2"]
    t2, t3, t4, t5 = K["t2"], K["t3"], K["t4"], K["t5"]
    seps = d * separation_at_contrast(delta_mags, separations, contrasts)

    def fraction(seps):
        max_Porbs = ((4 * pi ** 2) / (G * M_s * Msun) * (seps * au) ** 3) ** (1 / 2) / 86400
        lp = np.log10(max_Porbs)
        t2_partial = 0.5 * (lp - 1.0) * (2.0 * f1 + slope * (lp - 1.0))
        t3_partial = 0.5 * alpha * (lp ** 2 - 5.4 * lp + 6.8) + f2 * (lp - 2.0)
        t4_partial = (alpha * dlogP * (lp - 3.4) + f2 * (lp - 3.4)
                      + slope2 * (0.238095 * lp ** 2 - 0.952381 * lp + 0.485714))
        t5_partial = f3 * (3.33333 - 17.3566 * np.exp(-0.3 * lp))
        f_comp = np.zeros(len(seps))
        seg2 = (lp >= 1.0) & (lp < 2.0)
        seg3 = (lp >= 2.0) & (lp < 3.4)
        seg4 = (lp >= 3.4) & (lp < 5.5)
        seg5 = (lp >= 5.5) & (lp < 8.0)
        seg6 = lp >= 8.0
        if first_decade:
            f_comp[seg2] = t2_partial[seg2]
            f_comp[seg3] = t2 + t3_partial[seg3]
            f_comp[seg4] = t2 + t3 + t4_partial[seg4]
            f_comp[seg5] = t2 + t3 + t4 + t5_partial[seg5]
            f_comp[seg6] = t2 + t3 + t4 + t5
        else:
            f_comp[seg4] = t4_partial[seg4]
            f_comp[seg5] = t4 + t5_partial[seg5]
            f_comp[seg6] = t4 + t5
        if M_act >= 1.0:
            return np.log(f_comp)
        f_act = 0.65 * f_comp + 0.35 * f_comp * M_act
        f_act[f_act < 0.0] = 0.0
        return np.log(f_act)

    return _hostpar.pmap_concat(fraction, seps.shape[0], seps)


def lnprior_bound_TP(M_s, plx, delta_mags, separations, contrasts):
    """Bound-companion prior for planet scenarios: companion period > 10^3.4 d (priors.py:580-782)."""
    return _bound_companion_lnprior(M_s, plx, delta_mags, separations, contrasts, False)


def lnprior_bound_EB(M_s, plx, delta_mags, separations, contrasts):
    """Bound-companion prior for EB scenarios: tertiary period > 10 d (priors.py:784-984)."""
    return _bound_companion_lnprior(M_s, plx, delta_mags, separations, contrasts, True)


def lnprior_background(N_comp, delta_mags, separations, contrasts):
    """ln probability of a chance-aligned background star inside the contrast-curve limit
    (natural log; priors.py:986-1005)."""
    seps = separation_at_contrast(delta_mags, separations, contrasts)
    return np.log((N_comp / 0.1) * (1 / 3600) ** 2 * seps ** 2)
