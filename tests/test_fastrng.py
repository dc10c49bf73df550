"""The bulk continuation of numpy's global generator (triceratops_b200/_fastrng.py over
csrc/host_rng.c) must be indistinguishable from numpy / scipy: same numbers, same generator
state afterwards -- the prior draws of the lnZ_* functions depend on it (reference
marginal_likelihoods.py:101-104 and priors.py:134-155 draw from np.random)."""
import numpy as np
import pytest
from scipy.stats import powerlaw

import _workloads
from triceratops_b200 import _fastrng


@pytest.fixture(autouse=True)
def _needs_helper():
    if _fastrng._load() is None:
        pytest.fail("libtriceratops_host.so missing or its self-check failed")


def _same_state():
    a = np.random.get_state()
    return a[1].copy(), a[2], a[3], a[4]


@pytest.mark.parametrize("seed", [0, 1, 2026])
@pytest.mark.parametrize("n", [4096, 4097, 10_000, 312 * 40, 1_000_003])
def test_rand_skip_uniform_randint_powerlaw(seed, n):
    np.random.seed(seed)
    np.random.rand(3)                     # an odd position inside the first block
    want = [np.random.rand(n), np.random.rand(n)[:0], np.random.uniform(2.5, 3.5, n),
            np.random.randint(0, 2499, n), np.random.randint(0, 2500, n),
            powerlaw.rvs(0.2, size=n), powerlaw.rvs(0.6, size=n), np.random.rand(n)]
    want_state = _same_state()
    np.random.seed(seed)
    np.random.rand(3)
    got = [_fastrng.rand(n)]
    _fastrng.skip(n)
    got += [np.empty(0), _fastrng.uniform(2.5, 3.5, n), _fastrng.randint(0, 2499, n),
            _fastrng.randint(0, 2500, n), _fastrng.powerlaw_rvs(0.2, n),
            _fastrng.powerlaw_rvs(0.6, n), _fastrng.rand(n)]
    for w, g in zip(want, got):
        assert w.dtype == g.dtype and np.array_equal(w, g)
    got_state = _same_state()
    # the same stream position: whatever is drawn next is identical
    assert np.array_equal(np.random.rand(1000), (np.random.set_state(
        ("MT19937",) + want_state) or np.random.rand(1000)))
    assert got_state[2:] == want_state[2:]


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [4096, 40_000, 300_007])
def test_beta_rvs_is_scipys(seed, n):
    """The chunk-parallel walk of numpy's legacy_beta: values, stream position and the cached
    gaussian left in the state (odd seeds enter with one cached)."""
    from scipy.stats import beta
    out = []
    for fn in (lambda: beta.rvs(0.867, 3.030, size=n), lambda: _fastrng.beta_rvs(0.867, 3.030, n)):
        np.random.seed(seed)
        np.random.rand(3)
        if seed % 2:
            np.random.standard_normal(1)
        out.append((fn(), np.random.rand(5), np.random.standard_normal(3)))
    for a, b in zip(*out):
        assert np.array_equal(a, b)


def test_gaussian_cache_survives():
    """has_gauss / cached_gaussian of the legacy state pass through untouched."""
    np.random.seed(5)
    np.random.standard_normal(1)          # leaves a cached gaussian behind
    a = (np.random.rand(5000), np.random.standard_normal(2))
    np.random.seed(5)
    np.random.standard_normal(1)
    b = (_fastrng.rand(5000), np.random.standard_normal(2))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_engine_calls_are_identical_with_and_without_the_helper(monkeypatch):
    """The columns calc_probs hands to the engine (all 12 calls, N = 20000) with the bulk
    generator and with numpy's own calls."""
    lc = _workloads.lightcurve(2)
    fast, _ = _workloads.record_calls(2, 20_000, 11, lc)
    monkeypatch.setattr(_fastrng, "_lib", None)
    monkeypatch.setattr(_fastrng, "_checked", True)
    slow, _ = _workloads.record_calls(2, 20_000, 11, lc)

    def key(c):
        return (c["kind"], c["is_host"], float(np.sum(c["cols"]["inc"])))
    for a, b in zip(sorted(fast, key=key), sorted(slow, key=key)):
        assert a["kind"] == b["kind"] and a["is_host"] == b["is_host"]
        for k in a["cols"]:
            if a["cols"][k] is None:
                assert b["cols"][k] is None
            else:
                assert np.array_equal(np.asarray(a["cols"][k]), np.asarray(b["cols"][k])), k
        assert (a["extra_mask"] is None) == (b["extra_mask"] is None)
        if a["extra_mask"] is not None:
            assert np.array_equal(a["extra_mask"], b["extra_mask"])


@pytest.mark.parametrize("margin", [64, -3000, -10_000_000])
@pytest.mark.parametrize("hi", [2, 3, 1000, 2499, 65536, 2 ** 31 + 5])
def test_randint_bulk_matches_numpy_also_when_the_word_estimate_falls_short(margin, hi):
    """randint produces its words in bulk from an estimate of the rejection rate; a short
    estimate (forced here) must refill and land on the same values and stream position."""
    L = _fastrng._load()
    n = 50_001
    np.random.seed(5)
    np.random.rand(7)
    want, after = np.random.randint(0, hi, n), np.random.rand(5)
    np.random.seed(5)
    np.random.rand(7)
    L.trih_debug_randint_margin.argtypes = [__import__("ctypes").c_int64]
    L.trih_debug_randint_margin(margin)
    try:
        got = _fastrng.randint(0, hi, n)
    finally:
        L.trih_debug_randint_margin(64)
    assert got.dtype == want.dtype and np.array_equal(got, want)
    assert np.array_equal(np.random.rand(5), after)
