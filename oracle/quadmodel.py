"""
ORACLE (test infrastructure, not product code) -- restated quadratic transit model.

PARITY UNPINNED.  The arithmetic of the reference's hot loop lives in a third-party
package that is absent from /root/reference and from this image:

    pytransit == 2.2   (reference setup.py:25), class ``QuadraticModel``

Reference call sites (``triceratops/likelihoods.py``): import + model instances
:15, :24-25; scalar path :61-71, :120-145; vectorised path :346-349, :412-415,
:420-423.  No reference test and no golden vector pins numbers at that boundary
(SURVEY.md section 4 / 8c), and the package cannot be run here, so this file
*defines* the semantics the CUDA kernels are held to.  It restates the published
algorithms PyTransit 2.x uses for ``QuadraticModel(interpolate=False)``:

  * orbit: mean anomaly -> true anomaly by bilinear interpolation in a pre-computed
    (eccentricity x mean anomaly) table of ``f - M`` built with a Newton solve,
    then the projected star-planet separation z (PyTransit ``orbits_py``:
    ``mean_anomaly_offset``, ``ta_ip_calculate_table``, ``ta_ip_s``, ``z_from_ta_s``);
  * occultation: Mandel & Agol (2002) quadratic limb darkening, Table 3 cases,
    with the Hastings polynomial approximations of K and E (Abramowitz & Stegun
    17.3.34 / 17.3.36) and Bulirsch's iteration for the complete elliptic integral
    of the third kind (PyTransit ``ma_quadratic_nb``: ``eval_quad_z_s``, ``ellk``,
    ``ellec``, ``ellpicb``);
  * exposure supersampling: mean over ``nsamples`` sub-exposures at offsets
    ``exptime * ((i - 0.5)/nsamples - 0.5)``, i = 1..nsamples.

Choices made where the third-party behaviour is undefined (out-of-bounds table
reads in numba): the table indices are clamped to the last cell, i.e. the last
cell is linearly extrapolated for e >= 0.95 or M == pi.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Independent known-answer checks of the mathematics
(closed-form uniform source, brute-force quadrature, scipy elliptic integrals) are
in tests/test_oracle_model.py.
"""
import math

import numpy as np

try:  # numba only accelerates the oracle; pure Python gives the same numbers
    from numba import njit
except Exception:  # pragma: no cover
    def njit(*a, **k):
        def wrap(f):
            return f
        if len(a) == 1 and callable(a[0]):
            return a[0]
        return wrap

HALF_PI = 0.5 * math.pi
TWO_PI = 2.0 * math.pi
INV_PI = 1.0 / math.pi

# table geometry (PyTransit ``ta_ip_calculate_table`` defaults)
TABLE_NE = 256
TABLE_NM = 512
TABLE_MAX_E = 0.95


# --------------------------------------------------------------------------- orbit
@njit(cache=True)
def ta_newton(ma, e):
    """True anomaly from mean anomaly by Newton iteration on Kepler's equation."""
    ea = ma
    err = 0.05
    k = 0
    while abs(err) > 1e-8 and k < 1000:
        err = ea - e * math.sin(ea) - ma
        ea = ea - err / (1.0 - e * math.cos(ea))
        k += 1
    sta = math.sqrt(1.0 - e * e) * math.sin(ea) / (1.0 - e * math.cos(ea))
    cta = (math.cos(ea) - e) / (1.0 - e * math.cos(ea))
    return math.atan2(sta, cta)


@njit(cache=True)
def make_orbit_table(ne=TABLE_NE, nm=TABLE_NM, max_e=TABLE_MAX_E):
    """(es, ms, tae) with tae[i, j] = f(ms[j], es[i]) - ms[j] on [0, max_e] x [0, pi]."""
    es = np.linspace(0.0, max_e, ne)
    ms = np.linspace(0.0, math.pi, nm)
    tae = np.zeros((ne, nm))
    for i in range(ne):
        for j in range(nm):
            tae[i, j] = ta_newton(ms[j], es[i]) - ms[j]
    return es, ms, tae


@njit(cache=True)
def mean_anomaly_offset(e, w):
    """Mean anomaly at mid-transit (true anomaly pi/2 - w)."""
    off = math.atan2(math.sqrt(1.0 - e * e) * math.sin(HALF_PI - w),
                     e + math.cos(HALF_PI - w))
    off -= e * math.sin(off)
    return off


@njit(cache=True)
def ta_ip(t, t0, p, e, w, es, ms, tae):
    """True anomaly at time t via the bilinear (e, M) table."""
    ne = es.size
    nm = ms.size
    de = es[1] - es[0]
    dm = ms[1] - ms[0]

    ie = int(math.floor(e / de))
    if ie > ne - 2:
        ie = ne - 2
    ae = (e - de * ie) / de

    off = mean_anomaly_offset(e, w)
    ma = (TWO_PI * (t - (t0 - off * p / TWO_PI)) / p) % TWO_PI
    if ma < math.pi:
        x = ma
        s = 1.0
    else:
        x = TWO_PI - ma
        s = -1.0
    im = int(math.floor(x / dm))
    if im > nm - 2:
        im = nm - 2
    am = (x - im * dm) / dm

    d = (tae[ie, im] * (1.0 - ae) * (1.0 - am)
         + tae[ie + 1, im] * ae * (1.0 - am)
         + tae[ie, im + 1] * (1.0 - ae) * am
         + tae[ie + 1, im + 1] * ae * am)
    return ma + s * d


@njit(cache=True)
def z_from_ta(ta, a, i, e, w):
    """Projected separation [stellar radii]; negative on the far side of the orbit."""
    swt = math.sin(w + ta)
    si = math.sin(i)
    z = a * (1.0 - e * e) / (1.0 + e * math.cos(ta)) * math.sqrt(1.0 - swt * swt * si * si)
    if swt < 0.0:
        z = -z
    return z


@njit(cache=True)
def z_ip(t, t0, p, a, i, e, w, es, ms, tae):
    return z_from_ta(ta_ip(t, t0, p, e, w, es, ms, tae), a, i, e, w)


# ------------------------------------------------------------- elliptic integrals
@njit(cache=True)
def ellk(k):
    """Complete elliptic integral of the first kind, Hastings polynomial (A&S 17.3.34)."""
    m1 = 1.0 - k * k
    ek1 = 1.38629436112 + m1 * (0.09666344259 + m1 * (0.03590092383
          + m1 * (0.03742563713 + m1 * 0.01451196212)))
    ek2 = (0.5 + m1 * (0.12498593597 + m1 * (0.06880248576
          + m1 * (0.03328355346 + m1 * 0.00441787012)))) * math.log(m1)
    return ek1 - ek2


@njit(cache=True)
def ellec(k):
    """Complete elliptic integral of the second kind, Hastings polynomial (A&S 17.3.36)."""
    m1 = 1.0 - k * k
    ee1 = 1.0 + m1 * (0.44325141463 + m1 * (0.0626060122
          + m1 * (0.04757383546 + m1 * 0.01736506451)))
    ee2 = m1 * (0.2499836831 + m1 * (0.09200180037 + m1 * (0.04069697526
          + m1 * 0.00526449639))) * math.log(1.0 / m1)
    return ee1 + ee2


@njit(cache=True)
def ellpicb(n, k):
    """Complete elliptic integral of the third kind, Bulirsch (1965) iteration.

    Evaluates the integral of dθ / ((1 + n sin²θ) sqrt(1 - k² sin²θ)) over [0, π/2].
    """
    kc = math.sqrt(1.0 - k * k)
    e = kc
    p = math.sqrt(n + 1.0)
    m0 = 1.0
    c = 1.0
    d = 1.0 / p
    for _ in range(1000):
        f = c
        c = d / p + c
        g = e / p
        d = 2.0 * (f * g + d)
        p = g + p
        g = m0
        m0 = kc + m0
        if abs(1.0 - kc / g) > 1e-8:
            kc = 2.0 * math.sqrt(e)
            e = kc * m0
        else:
            return HALF_PI * (c * m0 + d) / (m0 * (m0 + p))
    return 0.0


# ------------------------------------------------------------------- occultation
@njit(cache=True)
def eval_quad(z, k, u1, u2):
    """Relative flux of a quadratically limb-darkened star occulted at separation z."""
    if abs(z - k) < 1e-6:
        z += 1e-6
    if z > 1.0 + k or z < 0.0:
        return 1.0
    if k >= 1.0 and z <= k - 1.0:
        return 0.0

    omega = 1.0 - u1 / 3.0 - u2 / 6.0
    k2 = k * k
    z2 = z * z
    x1 = (k - z) ** 2
    x2 = (k + z) ** 2
    x3 = k * k - z * z
    le = 0.0
    ld = 0.0
    ed = 0.0
    kap0 = 0.0
    kap1 = 0.0

    # uniform-source term
    if z >= abs(1.0 - k) and z <= 1.0 + k:
        kap1 = math.acos(min((1.0 - k2 + z2) / 2.0 / z, 1.0))
        kap0 = math.acos(min((k2 + z2 - 1.0) / 2.0 / k / z, 1.0))
        le = k2 * kap0 + kap1
        le = (le - 0.5 * math.sqrt(max(4.0 * z2 - (1.0 + z2 - k2) ** 2, 0.0))) * INV_PI
    if z <= 1.0 - k:
        le = k2

    # limb-darkening terms, Mandel & Agol (2002) Table 3
    if abs(z - k) < 1e-4 * (z + k):
        # edge of the occultor at the centre of the disc
        if k == 0.5:
            ld = 1.0 / 3.0 - 4.0 * INV_PI / 9.0
            ed = 3.0 / 32.0
        elif z > 0.5:
            q = 0.5 / k
            Kk = ellk(q)
            Ek = ellec(q)
            ld = (1.0 / 3.0 + 16.0 * k / 9.0 * INV_PI * (2.0 * k2 - 1.0) * Ek
                  - (32.0 * k ** 4 - 20.0 * k2 + 3.0) / 9.0 * INV_PI / k * Kk)
            ed = 1.0 / 2.0 * INV_PI * (kap1 + k2 * (k2 + 2.0 * z2) * kap0
                 - (1.0 + 5.0 * k2 + z2) / 4.0 * math.sqrt((1.0 - x1) * (x2 - 1.0)))
        else:
            q = 2.0 * k
            Kk = ellk(q)
            Ek = ellec(q)
            ld = 1.0 / 3.0 + 2.0 / 9.0 * INV_PI * (4.0 * (2.0 * k2 - 1.0) * Ek
                 + (1.0 - 4.0 * k2) * Kk)
            ed = k2 / 2.0 * (k2 + 2.0 * z2)
    elif ((z > 0.5 + abs(k - 0.5) and z < 1.0 + k)
          or (k > 0.5 and z > abs(1.0 - k) * 1.0001 and z < k)):
        # occultor crosses the limb (case III)
        q = math.sqrt((1.0 - x1) / 4.0 / z / k)
        Kk = ellk(q)
        Ek = ellec(q)
        n = 1.0 / x1 - 1.0
        Pk = ellpicb(n, q)
        ld = (1.0 / 9.0 * INV_PI / math.sqrt(k * z)
              * (((1.0 - x2) * (2.0 * x2 + x1 - 3.0) - 3.0 * x3 * (x2 - 2.0)) * Kk
                 + 4.0 * k * z * (z2 + 7.0 * k2 - 4.0) * Ek - 3.0 * x3 / x1 * Pk))
        if z < k:
            ld += 2.0 / 3.0
        ed = 1.0 / 2.0 * INV_PI * (kap1 + k2 * (k2 + 2.0 * z2) * kap0
             - (1.0 + 5.0 * k2 + z2) / 4.0 * math.sqrt((1.0 - x1) * (x2 - 1.0)))
    elif k <= 1.0 and z < (1.0 - k) * 1.0001:
        # occultor inside the disc (case IV)
        q = math.sqrt((x2 - x1) / (1.0 - x1))
        Kk = ellk(q)
        Ek = ellec(q)
        n = x2 / x1 - 1.0
        Pk = ellpicb(n, q)
        ld = (2.0 / 9.0 * INV_PI / math.sqrt(1.0 - x1)
              * ((1.0 - 5.0 * z2 + k2 + x3 * x3) * Kk
                 + (1.0 - x1) * (z2 + 7.0 * k2 - 4.0) * Ek - 3.0 * x3 / x1 * Pk))
        if z < k:
            ld += 2.0 / 3.0
        if abs(k + z - 1.0) < 1e-4:
            ld = (2.0 / 3.0 * INV_PI * math.acos(1.0 - 2.0 * k)
                  - 4.0 / 9.0 * INV_PI * math.sqrt(k * (1.0 - k)) * (3.0 + 2.0 * k - 8.0 * k2))
        ed = k2 / 2.0 * (k2 + 2.0 * z2)

    return 1.0 - ((1.0 - u1 - 2.0 * u2) * le + (u1 + 2.0 * u2) * ld + u2 * ed) / omega


# ------------------------------------------------------------------- light curves
@njit(cache=True)
def _model_pv(time, pvp, ldc, exptime, nsamples, es, ms, tae):
    npv = pvp.shape[0]
    npt = time.size
    flux = np.zeros((npv, npt))
    for ipv in range(npv):
        k = pvp[ipv, 0]
        t0 = pvp[ipv, 1]
        p = pvp[ipv, 2]
        a = pvp[ipv, 3]
        inc = pvp[ipv, 4]
        e = pvp[ipv, 5]
        w = pvp[ipv, 6]
        u1 = ldc[ipv, 0]
        u2 = ldc[ipv, 1]
        for j in range(npt):
            acc = 0.0
            for isample in range(1, nsamples + 1):
                toff = exptime * ((isample - 0.5) / nsamples - 0.5)
                z = z_ip(time[j] + toff, t0, p, a, inc, e, w, es, ms, tae)
                if z > 1.0 + k:
                    acc += 1.0
                else:
                    acc += eval_quad(z, k, u1, u2)
            flux[ipv, j] = acc / nsamples
    return flux


_TABLE = None


def orbit_table():
    global _TABLE
    if _TABLE is None:
        _TABLE = make_orbit_table()
    return _TABLE


class QuadraticModel:
    """Duck-type of the subset of ``pytransit.QuadraticModel`` the reference calls."""

    def __init__(self, interpolate=False, **kwargs):
        if interpolate:
            raise NotImplementedError("the reference only uses interpolate=False")
        self.time = None
        self.exptime = 0.0
        self.nsamples = 1
        self._es, self._ms, self._tae = orbit_table()

    def set_data(self, time, lcids=None, pbids=None, nsamples=None, exptimes=None, epids=None):
        self.time = np.ascontiguousarray(time, dtype=np.float64)
        self.nsamples = 1 if nsamples is None else int(np.ravel(nsamples)[0])
        self.exptime = 0.0 if exptimes is None else float(np.ravel(exptimes)[0])

    def evaluate_pv(self, pvp, ldc, copy=True):
        pvp = np.atleast_2d(np.asarray(pvp, dtype=np.float64))
        ldc = np.atleast_2d(np.asarray(ldc, dtype=np.float64))
        flux = _model_pv(self.time, np.ascontiguousarray(pvp), np.ascontiguousarray(ldc),
                         self.exptime, self.nsamples, self._es, self._ms, self._tae)
        return np.squeeze(flux)

    def evaluate_ps(self, k, ldc, t0, p, a, i, e=0.0, w=0.0, copy=True):
        pvp = np.array([[k, t0, p, a, i, e, w]], dtype=np.float64)
        return self.evaluate_pv(pvp, np.asarray(ldc, dtype=np.float64).reshape(1, 2))
