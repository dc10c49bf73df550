"""triceratops_b200: B200 (sm_100a) engine for the marginal-likelihood path of TRICERATOPS.

Drop-in for `target.calc_probs(...)` and the ten `lnZ_*` functions of the reference
(stevengiacalone/triceratops, triceratops/marginal_likelihoods.py): the host code stays in Python
and mirrors the reference signatures; geometry, light curves, chi^2 and the log-mean-exp run in
hand-written CUDA kernels behind a C ABI (include/triceratops_b200.h).  No CPU fallback.
"""
__version__ = "0.1.0"
