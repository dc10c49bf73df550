"""Once per process: csrc/host_blocks.c against the numpy statements it mirrors, on a small
instance that goes through every one of numpy's inner loops the C code borrows (power with a
scalar exponent, with a scalar base and through the `**` fast paths, log10, log, exp, arccos).
Any differing bit disables the C path for the process (tests/test_host_blocks.py is the
exhaustive comparison, over every scenario and option)."""
import numpy as np


def _numpy_peb(M_s, R_s, Teff, plx, P_mean, c_comp, x_inc, x_q, x_e, x_w):
    from . import marginal_likelihoods as ml
    from .funcs import stellar_relations
    from .priors import lnprior_bound_EB, sample_inc, sample_q, sample_q_companion, sample_w
    qs_comp = sample_q_companion(c_comp, M_s)
    ml._ecc_binary(x_e, P_mean)
    incs, qs, argps = sample_inc(x_inc), sample_q(x_q, M_s), sample_w(x_w)
    masses = qs * M_s
    radii, _ = stellar_relations(masses, np.full(len(qs), R_s), np.full(len(qs), Teff))
    fluxratios = ml._fluxratio(masses, M_s)
    fluxratios_comp = ml._fluxratio(qs_comp * M_s, M_s)
    lnprior = ml._bound_prior(lnprior_bound_EB, M_s, plx, len(qs), None, None,
                              fluxratios_comp / (1 - fluxratios_comp), None)
    return (incs, qs, argps, masses, radii, fluxratios, M_s + masses, fluxratios_comp, lnprior,
            qs_comp != 0.0), x_e


def self_check():
    from . import _blocks, _hostpar
    rng = np.random.default_rng(20260117)
    n = 20011
    for M_s, P_mean in ((0.93, 4.2), (1.31, 17.0), (0.24, 4.2)):
        dev = [rng.random(n) for _ in range(5)]
        dev[0][:7] = (0.0, 1.0 - 2.0 ** -53, 0.5, 1e-300, 0.9999999, 0.3, 0.95)
        a = [d.copy() for d in dev]
        b = [d.copy() for d in dev]
        _hostpar._tl.inline = True         # (the numpy side: plain calls, no thread pool)
        try:
            want, want_e = _numpy_peb(M_s, 0.95, 5400.0, 8.1, P_mean, *a)
        finally:
            _hostpar._tl.inline = False
        got = _blocks.run("PEB", n, M_s=M_s, R_s=0.95, Teff=5400.0, c_comp=b[0], x_inc=b[1],
                          x_q=b[2], x_e=b[3], x_w=b[4], P_mean=P_mean, plx=8.1, bound_kind="EB",
                          _force=True)
        if got is None or len(got) != len(want):
            return False
        for w, g in zip(want, got):
            if w.dtype != g.dtype or not np.array_equal(w, g, equal_nan=True):
                return False
        if not np.array_equal(want_e, b[3], equal_nan=True):
            return False
    return True
