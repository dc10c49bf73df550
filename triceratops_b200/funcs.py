"""Host-side stellar relations and table readers that feed the marginal-likelihood path.

Numerics part of the reference's triceratops/funcs.py (the network / catalogue half is out of
scope).  Every function returns bit-identical arrays to its
reference counterpart for the same inputs: same spline nodes, same scipy objects, same
operation order.
"""
import os

import numpy as np
from pandas import read_csv
from scipy.interpolate import InterpolatedUnivariateSpline as _Spline

from . import _hostpar
from ._hostpar import splev as _splev

# --- mass -> (radius, Teff): Torres (2010) above 0.63 Msun, cool-dwarf relation below
#     (funcs.py:19-51)
_HOT_M = np.array([0.26, 0.47, 0.59, 0.69, 0.87, 0.98, 1.085, 1.4, 1.65, 2.0, 2.5, 3.0, 4.4,
                   15.0, 40.0])
_HOT_T = np.array([3170, 3520, 3840, 4410, 5150, 5560, 5940, 6650, 7300, 8180, 9790, 11400,
                   15200, 30000, 42000])
_HOT_R = np.array([0.28, 0.47, 0.60, 0.72, 0.9, 1.05, 1.2, 1.55, 1.8, 2.1, 2.4, 2.6, 3.0, 6.2,
                   11.0])
_COOL_M = np.array([0.1, 0.135, 0.2, 0.35, 0.48, 0.58, 0.63])
_COOL_T = np.array([2800, 3000, 3200, 3400, 3600, 3800, 4000])
_COOL_R = np.array([0.12, 0.165, 0.23, 0.36, 0.48, 0.585, 0.6])
_hot_T, _hot_R = _Spline(_HOT_M, _HOT_T), _Spline(_HOT_M, _HOT_R)
_cool_T, _cool_R = _Spline(_COOL_M, _COOL_T), _Spline(_COOL_M, _COOL_R)

# --- mass -> log10 flux relative to a ~1 Msun star, per band (funcs.py:81-119)
_FLUX_SPLINES = {
    "TESS": _Spline(np.array([0.1, 0.15, 0.23, 0.4, 0.58, 0.7, 0.9, 1.15, 1.45, 2.2, 2.8]),
                    np.array([-3, -2.5, -2, -1.5, -1, -0.5, 0, 0.5, 1, 1.5, 2])),
    "J": _Spline(np.array([0.1, 0.2, 0.5, 0.75, 1.0, 1.5, 2.0, 2.5, 3]),
                 np.array([-5.7, -3.8, -1.6, 0, 1.2, 2.9, 3.3, 4, 6]) / 2.5),
    "H": _Spline(np.array([0.1, 0.23, 0.5, 0.75, 1.0, 1.5, 2.0, 2.5, 3]),
                 np.array([-4.9, -2.8, -0.9, 0.6, 1.5, 3, 3.3, 4, 6]) / 2.5),
    "K": _Spline(np.array([0.1, 0.2, 0.35, 0.5, 0.75, 1.0, 1.5, 2.0, 2.5, 3]),
                 np.array([-4.7, -2.9, -1.7, -0.7, 0.6, 1.6, 3, 3.3, 4, 6]) / 2.5),
}
_FLUX_SPLINES["Vis"] = _FLUX_SPLINES["TESS"]


def stellar_relations(Masses, max_Radii, max_Teffs):
    """Radii [Rsun] and Teffs [K] for `Masses` [Msun], capped at (max_Radii, max_Teffs) and
    floored at (0.1, 2800) -- reference funcs.py:54-79."""
    Masses = np.asarray(Masses)
    hot = Masses > 0.63
    cool = Masses <= 0.63
    Radii = np.zeros(len(Masses))
    Teffs = np.zeros(len(Masses))
    m_hot, m_cool = Masses[hot], Masses[cool]
    Radii[hot] = _splev(_hot_R, m_hot)
    Teffs[hot] = _splev(_hot_T, m_hot)
    Radii[cool] = _splev(_cool_R, m_cool)
    Teffs[cool] = _splev(_cool_T, m_cool)
    big = Radii > max_Radii
    Radii[big] = max_Radii[big]
    warm = Teffs > max_Teffs
    Teffs[warm] = max_Teffs[warm]
    Radii[Radii < 0.1] = 0.1
    Teffs[Teffs < 2800] = 2800
    return Radii, Teffs


def flux_relation(Masses, filt: str = "TESS"):
    """Flux relative to a ~1 Msun star in band `filt` (TESS, Vis, J, H, K) -- funcs.py:121-140."""
    y = _splev(_FLUX_SPLINES[filt], Masses)
    return _hostpar.pmap_concat(lambda v: 10 ** v, y.shape[0], y) if y.ndim else 10 ** y


def renorm_flux(flux, flux_err, star_fluxratio: float):
    """Light curve as it would look if only this star were in the aperture (funcs.py:164-177)."""
    return (flux - (1 - star_fluxratio)) / star_fluxratio, flux_err / star_fluxratio


_contrast_cache = {}


def file_to_contrast_curve(contrast_curve_file: str):
    """(separations [arcsec], |contrasts| [mag]) from a two-column CSV (funcs.py:203-219)."""
    st = os.stat(contrast_curve_file)
    key = (os.path.abspath(contrast_curve_file), st.st_mtime_ns, st.st_size)
    if key not in _contrast_cache:      # (every chunk of every P*/S*/D*/B* scenario asks)
        if len(_contrast_cache) > 8:
            _contrast_cache.clear()
        data = np.loadtxt(contrast_curve_file, delimiter=',')
        _contrast_cache[key] = (data.T[0].copy(), np.abs(data.T[1]))
    return _contrast_cache[key]


def separation_at_contrast(delta_mags, separations, contrasts):
    """Separation beyond which a companion of contrast delta_mags is ruled out (funcs.py:222-238;
    np.interp is applied as-is even when the contrasts are not monotonic)."""
    return np.interp(delta_mags, contrasts, separations)


_trilegal_cache = {}


def _read_trilegal(fname):
    """The saved table minus its two trailer rows (funcs.py:353); parsed once per file version
    (a calc_probs reads it four times, a sweep once per target)."""
    st = os.stat(fname)
    key = (os.path.abspath(fname), st.st_mtime_ns, st.st_size)
    if key not in _trilegal_cache:
        if len(_trilegal_cache) > 8:
            _trilegal_cache.clear()
        _trilegal_cache[key] = read_csv(fname)[:-2]
    return _trilegal_cache[key]


def trilegal_results(trilegal_fname: str, Tmag: float):
    """Background-star population fainter than the target from a saved TRILEGAL table
    (funcs.py:335-403): (Tmags, Masses, loggs, Teffs, Zs, Jmags, Hmags, Kmags)."""
    df = _read_trilegal(trilegal_fname)
    Masses = df["Mact"].values
    loggs = df["logg"].values
    Teffs = 10 ** df["logTe"].values
    Zs = np.array(df["[M/H]"], dtype=float)
    Jmags = df["J"].values
    Hmags = df["H"].values
    Kmags = df["Ks"].values
    if "TESS" in df.columns:
        Tmags = df["TESS"].values
    else:
        # TRILEGAL v1.5 tables carry 2MASS only: T from J, Ks (Stassun et al. 2018, 2.2.1.1)
        c = Jmags - Kmags
        Tmags = np.zeros(df.shape[0])
        blue = (-0.1 <= c) & (c <= 0.70)
        red = (0.7 < c) & (c <= 1.0)
        Tmags[blue] = (Jmags[blue] + 1.22163 * c[blue] ** 3 - 1.74299 * c[blue] ** 2
                       + 1.89115 * c[blue] + 0.0563)
        Tmags[red] = (Jmags[red] - 269.372 * c[red] ** 3 + 668.453 * c[red] ** 2
                      - 545.64 * c[red] + 147.811)
        Tmags[c < -0.1] = Jmags[c < -0.1] + 0.5
        Tmags[c > 1.0] = Jmags[c > 1.0] + 1.75
    keep = Tmags >= Tmag
    return (Tmags[keep], Masses[keep], loggs[keep], Teffs[keep], Zs[keep], Jmags[keep],
            Hmags[keep], Kmags[keep])
