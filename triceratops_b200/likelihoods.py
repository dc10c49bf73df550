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

__all__ = ["simulate_TP_transit", "simulate_EB_transit", "simulate_TP_transit_p",
           "simulate_EB_transit_p", "lnL_TP_p", "lnL_EB_p", "lnL_EB_twin_p"]


def _prepare(time, flux, sigma, exptime, nsamples):
    eng = _dispatch.get_engine()
    eng.set_lightcurve(time, flux, sigma, exptime, nsamples)
    return eng


def _time_grid(time, exptime, nsamples):
    """Model evaluation only needs the time stamps: flux and sigma are placeholders."""
    time = np.ascontiguousarray(time, dtype=np.float64)
    return _prepare(time, np.ones_like(time), 1.0, exptime, nsamples), time


def simulate_TP_transit_p(time: np.ndarray, R_p: np.ndarray,
                          P_orb: float, inc: np.ndarray,
                          a: np.ndarray, R_s: np.ndarray,
                          u1: np.ndarray, u2: np.ndarray,
                          ecc: np.ndarray, argp: np.ndarray,
                          companion_fluxratio: np.ndarray,
                          companion_is_host: bool = False,
                          exptime: float = 0.00139,
                          nsamples: int = 20):
    """Diluted transiting-planet light curves, one row per draw (likelihoods.py:302-358).
    Unlike the reference (whose pytransit call squeezes), one draw still gives shape (1, npts)."""
    eng, time = _time_grid(time, exptime, nsamples)
    return eng.simulate_tp(time.size, R_p, P_orb, inc, a, R_s, u1, u2, ecc, argp,
                           companion_fluxratio, companion_is_host)


def simulate_EB_transit_p(time: np.ndarray, R_EB: np.ndarray,
                          EB_fluxratio: np.ndarray,
                          P_orb: float, inc: np.ndarray,
                          a: np.ndarray, R_s: np.ndarray,
                          u1: np.ndarray, u2: np.ndarray,
                          ecc: np.ndarray, argp: np.ndarray,
                          companion_fluxratio: np.ndarray,
                          companion_is_host: bool = False,
                          exptime: float = 0.00139,
                          nsamples: int = 20):
    """Diluted eclipsing-binary light curves and secondary-eclipse depths (likelihoods.py:361-439):
    returns (flux[n, npts], secdepth[n, 1])."""
    eng, time = _time_grid(time, exptime, nsamples)
    flux, sec = eng.simulate_eb(time.size, R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc,
                                argp, companion_fluxratio, companion_is_host)
    return flux, sec.reshape(-1, 1)


def simulate_TP_transit(time: np.ndarray, R_p: float, P_orb: float,
                        inc: float, a: float, R_s: float, u1: float,
                        u2: float, ecc: float, argp: float,
                        companion_fluxratio: float = 0.0,
                        companion_is_host: bool = False,
                        exptime: float = 0.00139,
                        nsamples: int = 20):
    """One transiting-planet light curve (likelihoods.py:27-80), as plot_fits uses it."""
    eng, time = _time_grid(time, exptime, nsamples)
    one = [np.array([float(x)]) for x in (R_p, P_orb, inc, a, R_s, u1, u2, ecc, argp,
                                          companion_fluxratio)]
    return eng.simulate_tp(time.size, *one, companion_is_host)[0]


def simulate_EB_transit(time: np.ndarray, R_EB: float,
                        EB_fluxratio: float, P_orb: float, inc: float,
                        a: float, R_s: float, u1: float, u2: float,
                        ecc: float, argp: float,
                        companion_fluxratio: float = 0.0,
                        companion_is_host: bool = False,
                        exptime: float = 0.00139,
                        nsamples: int = 20):
    """One eclipsing-binary light curve and its secondary depth (likelihoods.py:83-160).  The
    scalar function has its own radius-ratio rules (|k - 1| < 1e-6; secondary uses 1/k), kept."""
    eng, time = _time_grid(time, exptime, nsamples)
    one = [np.array([float(x)]) for x in (R_EB, EB_fluxratio, P_orb, inc, a, R_s, u1, u2, ecc,
                                          argp, companion_fluxratio)]
    flux, sec = eng.simulate_eb(time.size, *one, companion_is_host, scalar_rule=True)
    return flux[0], float(sec[0])


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
