"""Log-mean-exp and probability normalisation contracts: the known-answer cases of the
reference's tests/test_log_mean_exp.py (:24-270), replayed against the host implementation and
against the (max, scaled-sum) combine used across GPUs."""
import math

import numpy as np
import pytest
from scipy.special import logsumexp

from triceratops_b200._numerics import _log_mean_exp, _normalize_probabilities
from triceratops_b200.engine import combine_lse


def _parts(lnw, nsplit):
    """Per-rank (m, s, n_finite, n_posinf) records of contiguous slices, as the GPU returns."""
    out = []
    for chunk in np.array_split(np.asarray(lnw, float), nsplit):
        fin = np.isfinite(chunk)
        m = chunk[fin].max() if fin.any() else -math.inf
        s = float(np.exp(chunk[fin] - m).sum()) if fin.any() else 0.0
        out.append((m, s, int(fin.sum()), int(np.isposinf(chunk).sum())))
    return out


CASES = {
    "underflow": (-2000.0 * np.ones(100000), -2000.0),
    "mixed": (np.array([-1001., -1002., -np.inf, -np.inf, -1003., -np.inf, -1004., -np.inf,
                        -1005., -np.inf]),
              float(logsumexp([-1001., -1002., -1003., -1004., -1005.]) - np.log(10))),
    "denominator": (np.array([-1.0] + [-np.inf] * 9), -1.0 - np.log(10)),
    "all_neginf": (np.full(50, -np.inf), -np.inf),
    "single": (np.array([-3.5]), -3.5),
    "nan_is_zero_weight": (np.array([-10., np.nan, -11., np.nan]),
                           float(logsumexp([-10., -11.]) - np.log(4))),
    "posinf": (np.array([-10., np.inf, -11.]), np.inf),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_log_mean_exp_known_answers(name):
    lnw, want = CASES[name]
    got = _log_mean_exp(lnw, N_total=lnw.size)
    assert got == want or abs(got - want) < 1e-10


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("nsplit", [1, 2, 8])
def test_rank_combine_equals_log_mean_exp(name, nsplit):
    lnw, want = CASES[name]
    got = combine_lse(_parts(lnw, min(nsplit, lnw.size)), lnw.size)
    assert got == want or abs(got - want) < 1e-10


def test_legacy_shift_equivalence():
    rng = np.random.default_rng(0)
    lnL = rng.uniform(-650, -550, 5000)
    legacy = np.log(np.mean(np.nan_to_num(np.exp(lnL + 600)))) - 600
    assert abs(_log_mean_exp(lnL, N_total=lnL.size) - legacy) < 1e-12
    assert abs(combine_lse(_parts(lnL, 4), lnL.size) - legacy) < 1e-12


def test_size_mismatch_raises():
    with pytest.raises(ValueError):
        _log_mean_exp(np.zeros(5), N_total=4)


def test_normalize_probabilities_statuses():
    p, st = _normalize_probabilities(np.array([-10., -11., -np.inf]))
    assert st == "ok" and abs(p.sum() - 1) < 1e-12 and p[2] == 0
    p, st = _normalize_probabilities(np.full(3, -np.inf))
    assert st == "all_neginf" and not p.any()
    for bad in (np.nan, np.inf):
        p, st = _normalize_probabilities(np.array([-1., bad]))
        assert st == "anomaly" and not p.any()
