"""Stub of astropy.constants (oracle only): the five constants the reference reads as
``constants.X.cgs.value`` (likelihoods.py:17-21, marginal_likelihoods.py:13-17, priors.py:8-12,
funcs.py:12-16).  Values = astropy >= 4.0 defaults (CODATA 2018, IAU 2015 nominal), in cgs."""


class _Q:
    def __init__(self, v):
        self.value = v

    @property
    def cgs(self):
        return self


G = _Q(6.6743e-08)                 # cm^3 g^-1 s^-2
M_sun = _Q(1.988409870698051e+33)  # g   (GM_sun 1.3271244e20 m^3 s^-2 / G)
R_sun = _Q(69570000000.0)          # cm
R_earth = _Q(637810000.0)          # cm  (IAU 2015 nominal equatorial)
au = _Q(14959787070000.0)          # cm
