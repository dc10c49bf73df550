"""Physical constants in cgs, numerically equal to astropy >= 4.0 `constants.X.cgs.value`
(CODATA 2018 / IAU 2015 nominal), which the reference reads at likelihoods.py:17-21,
marginal_likelihoods.py:13-17, priors.py:8-12 and funcs.py:12-16.  The same literals are
compiled into the kernels (csrc/tri_model.cuh)."""
import numpy as np

G = 6.6743e-08
Msun = 1.988409870698051e+33
Rsun = 69570000000.0
Rearth = 637810000.0
au = 14959787070000.0
pi = np.pi
ln2pi = np.log(2 * pi)
