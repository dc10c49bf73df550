"""The per-draw likelihood seam of the reference (triceratops/likelihoods.py:443-587) on the GPU.

`lnL_TP_p`, `lnL_EB_p` and `lnL_EB_twin_p` keep the reference signatures and return convention:
arrays of +0.5*chi^2 over the already-masked draws (lnL_EB_p: +inf where the secondary eclipse
would have been detected, likelihoods.py:535-538).  The (n_draws x n_points) model matrix of the
reference (likelihoods.py:349, :486) is never materialised: each warp streams a draw's light
curve through registers.

Differences that follow from the reference's use of pytransit: none in the returned numbers
beyond rounding (see DESIGN.md "Parity"); `inc` is NOT converted to radians in place here
(likelihoods.py:344 mutates its argument, which is always a temporary in the reference).
"""
import numpy as np

from . import _dispatch


def _prepare(time, flux, sigma, exptime, nsamples):
    eng = _dispatch.get_engine()
    eng.set_lightcurve(time, flux, sigma, exptime, nsamples)
    return eng


def lnL_TP_p(time: np.ndarray, flux: np.ndarray, sigma: float,
             R_p: np.ndarray, P_orb: float, inc: np.ndarray,
             a: np.ndarray, R_s: np.ndarray,
             u1: np.ndarray, u2: np.ndarray,
             ecc: np.ndarray, argp: np.ndarray,
             companion_fluxratio: np.ndarray,
             companion_is_host: bool = False,
             exptime: float = 0.00139,
             nsamples: int = 20):
    """0.5*chi^2 of the transiting-planet model for each draw (likelihoods.py:443-487).
    Units: R_p [R_earth], P_orb [d], inc/argp [deg], a [cm], R_s [R_sun]."""
    eng = _prepare(time, flux, sigma, exptime, nsamples)
    return eng.lnl_tp(R_p, P_orb, inc, a, R_s, u1, u2, ecc, argp, companion_fluxratio,
                      companion_is_host)


def lnL_EB_p(time: np.ndarray, flux: np.ndarray, sigma: float,
             R_EB: np.ndarray, EB_fluxratio: np.ndarray,
             P_orb: float, inc: np.ndarray,
             a: np.ndarray, R_s: np.ndarray,
             u1: np.ndarray, u2: np.ndarray,
             ecc: np.ndarray, argp: np.ndarray,
             companion_fluxratio: np.ndarray,
             companion_is_host: bool = False,
             exptime: float = 0.00139,
             nsamples: int = 20):
    """0.5*chi^2 of the q < 0.95 eclipsing-binary model, +inf where secdepth >= 1.5 sigma
    (likelihoods.py:490-539).  R_EB [R_sun]."""
    eng = _prepare(time, flux, sigma, exptime, nsamples)
    return eng.lnl_eb(R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp,
                      companion_fluxratio, companion_is_host, twin=False)


def lnL_EB_twin_p(time: np.ndarray, flux: np.ndarray, sigma: float,
                  R_EB: np.ndarray, EB_fluxratio: np.ndarray,
                  P_orb: float, inc: np.ndarray,
                  a: np.ndarray, R_s: np.ndarray,
                  u1: np.ndarray, u2: np.ndarray,
                  ecc: np.ndarray, argp: np.ndarray,
                  companion_fluxratio: np.ndarray,
                  companion_is_host: bool = False,
                  exptime: float = 0.00139,
                  nsamples: int = 20):
    """0.5*chi^2 of the q >= 0.95, 2 x P_orb eclipsing-binary model (likelihoods.py:542-587);
    the caller passes the doubled period and the matching semi-major axis."""
    eng = _prepare(time, flux, sigma, exptime, nsamples)
    return eng.lnl_eb(R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc, argp,
                      companion_fluxratio, companion_is_host, twin=True)
