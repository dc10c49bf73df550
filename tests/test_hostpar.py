"""Host-side parallel helpers change WHERE the prior-draw transforms run, never their values:
chunked/threaded and C-spline results must be bit-identical to the serial numpy/scipy ones."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from triceratops_b200 import _hostpar, funcs, priors


@pytest.fixture()
def forced_chunks(monkeypatch):
    monkeypatch.setattr(_hostpar, "MIN_CHUNK", 777)
    monkeypatch.setattr(_hostpar, "N_THREADS", 4)
    monkeypatch.setattr(_hostpar, "_pool", ThreadPoolExecutor(4))


def _serial(monkeypatch):
    monkeypatch.setattr(_hostpar, "_pool", None)


def test_c_spline_is_bit_identical_to_scipy():
    assert _hostpar._host_lib() is not None, "libtriceratops_host.so not built"
    rng = np.random.default_rng(0)
    for spl in (funcs._hot_T, funcs._hot_R, funcs._cool_T, funcs._cool_R,
                *funcs._FLUX_SPLINES.values()):
        t = spl._eval_args[0]
        x = np.concatenate([rng.uniform(-0.5, 45, 50000), t, np.nextafter(t, np.inf),
                            np.nextafter(t, -np.inf), [0.0, -1.0, 1e3]])
        assert np.array_equal(_hostpar.splev(spl, x), spl(x))


def test_chunked_transforms_equal_serial(forced_chunks, monkeypatch, contrast_file):
    n = 50_003
    rng = np.random.default_rng(1)
    x = rng.random(n)
    Ms = rng.uniform(0.1, 1.5, n)
    m = rng.uniform(0.05, 3, n)
    dm = rng.uniform(0, 12, n)
    sep, con = funcs.file_to_contrast_curve(contrast_file)
    assert len(_hostpar.slices(n)) > 1

    def run():
        with np.errstate(divide="ignore"):
            return (priors.sample_rp(x.copy(), Ms, False), priors.sample_q(x.copy(), 0.811),
                    priors.sample_q_companion(x.copy(), 1.2), priors.sample_inc(x.copy()),
                    funcs.flux_relation(m, "K"),
                    *funcs.stellar_relations(m, np.full(n, 0.9), np.full(n, 5000.)),
                    priors.lnprior_bound_TP(0.811, 8.16, dm, sep, con),
                    priors.lnprior_bound_EB(1.3, 8.16, dm, sep, con))

    chunked = run()
    _serial(monkeypatch)
    assert len(_hostpar.slices(n)) == 1
    serial = run()
    for a, b in zip(chunked, serial):
        assert np.array_equal(a, b, equal_nan=True)
