"""Size-independent properties at BASELINE.json's full sizes (N = 1e6 draws, where the oracle
is too slow to be the checker): masks against the reference's numpy expressions, rank-combine
invariance, determinism, light-curve order invariance and prior linearity."""
import numpy as np
import pytest

from conftest import TOI465
from triceratops_b200._constants import G, Msun, Rearth, Rsun, pi
from triceratops_b200.engine import combine_lse

pytestmark = pytest.mark.gpu

N = 1_000_000


@pytest.fixture(scope="module")
def big(gpu_engine, toi465_lc):
    t, f, s = toi465_lc
    gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    np.random.seed(0)
    from triceratops_b200 import priors
    rps = priors.sample_rp(np.random.rand(N), np.full(N, TOI465["M"]), False)
    incs = priors.sample_inc(np.random.rand(N))
    eccs = priors.sample_ecc(np.random.rand(N), True, TOI465["P"])
    argps = priors.sample_w(np.random.rand(N))
    args = (rps, TOI465["P"], incs, eccs, argps, TOI465["M"], TOI465["R"], 0.4338, 0.2008, 0.0)
    return args, gpu_engine.eval_tp(N, *args, want_mask=True)


def test_mask_equals_reference_expressions_at_full_size(big):
    (rps, P, incs, eccs, argps, M, R, *_), r = big
    a = ((G * M * Msun) / (4 * pi ** 2) * (np.full(N, P) * 86400) ** 2) ** (1 / 3)
    e_corr = (1 + eccs * np.sin(argps * pi / 180)) / (1 - eccs ** 2)
    Ptra = (rps * Rearth + R * Rsun) / a * e_corr
    coll = (rps * Rearth + R * Rsun) > a * (1 - eccs)
    inc_min = np.full(N, 90.)
    inc_min[Ptra <= 1.] = np.arccos(Ptra[Ptra <= 1.]) * 180. / pi
    mask = (incs >= inc_min) & (coll == False)  # noqa: E712   marginal_likelihoods.py:107-123
    assert np.array_equal(mask, r.mask)
    assert np.array_equal(np.isfinite(r.lnL), mask)
    assert r.n_pass == mask.sum() and 0.05 * N < r.n_pass < 0.2 * N


def test_lnz_is_the_log_mean_exp_of_the_returned_lnl(big):
    from triceratops_b200._numerics import _log_mean_exp
    _, r = big
    assert abs(r.lnZ - _log_mean_exp(r.lnL, N_total=N)) < 1e-9


def test_sharded_evaluation_combines_to_the_same_lnz(big, gpu_engine):
    """What 2/4/8 ranks would compute: disjoint slices + the (max, scaled-sum) combine."""
    args, whole = big
    for world in (2, 8):
        parts, lnl = [], []
        for r in range(world):
            lo, hi = (r * N) // world, ((r + 1) * N) // world
            sl = [x[lo:hi] if np.ndim(x) else x for x in args]
            res = gpu_engine.eval_tp(hi - lo, *sl)
            parts.append((res.m, res.s, res.n_finite, res.n_posinf))
            lnl.append(res.lnL)
        assert abs(combine_lse(parts, N) - whole.lnZ) < 1e-9
        assert np.array_equal(np.concatenate(lnl), whole.lnL)   # per-draw results do not move


def test_repeat_is_bit_identical(big, gpu_engine):
    args, whole = big
    again = gpu_engine.eval_tp(N, *args)
    assert np.array_equal(again.lnL, whole.lnL) and again.lnZ == whole.lnZ


def test_constant_prior_shifts_lnz(big, gpu_engine):
    args, whole = big
    shifted = gpu_engine.eval_tp(N, *args, lnprior=np.full(N, -2.5), want_lnL=False)
    assert abs(shifted.lnZ - (whole.lnZ - 2.5)) < 1e-9
    nothing = gpu_engine.eval_tp(N, *args, lnprior=np.full(N, -np.inf), want_lnL=False)
    assert nothing.lnZ == -np.inf


def test_light_curve_order_does_not_matter(big, gpu_engine, toi465_lc):
    args, whole = big
    t, f, s = toi465_lc
    perm = np.random.default_rng(1).permutation(t.size)
    gpu_engine.set_lightcurve(t[perm], f[perm], s, 0.00139, 20)
    sub = [x[:50000] if np.ndim(x) else x for x in args]
    r = gpu_engine.eval_tp(50000, *sub)
    gpu_engine.set_lightcurve(t, f, s, 0.00139, 20)
    fin = np.isfinite(r.lnL)
    np.testing.assert_allclose(r.lnL[fin], whole.lnL[:50000][fin], rtol=1e-13)
