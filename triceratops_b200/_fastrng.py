"""numpy's global generator, continued in bulk (host-side prior draws only).

The reference draws every prior sample from `np.random` (MT19937 behind the legacy RandomState),
and parity "on identical host-drawn sample arrays (numpy seed)" needs that exact stream; at
N = 1e6 a calc_probs call consumes 57 x rand(N) and its wall time is the sequential generator.
csrc/host_rng.c continues the stream from `np.random.get_state()` with the vectorised MT19937
recurrence and hands the state back with `set_state()`:

    rand(n)                 == np.random.rand(n)
    skip(n)                 == np.random.rand(n) without materialising it (the reference draws
                               some arrays only for their length, e.g. priors.py:134-155)
    randint(low, high, n)   == np.random.randint(low, high, n)
    uniform(lo, hi, n)      == np.random.uniform(lo, hi, n)
    powerlaw_rvs(a, n)      == scipy.stats.powerlaw.rvs(a, size=n)
    beta_rvs(a, b, n)       == scipy.stats.beta.rvs(a, b, size=n)   (a < 1 < b: the planets'
                               eccentricity prior; numpy's rejection samplers walked in
                               parallel chunks that are stitched where their states coincide)

bit for bit.  The equality is verified against numpy itself the first time the module is used
(on a copy of the state); if the check fails or the helper library is missing, the functions
call numpy -- the values are the same either way, this is host-side preparation.
"""
import ctypes
import os

import numpy as np

from . import _build

_U32 = ctypes.POINTER(ctypes.c_uint32)
_D = ctypes.POINTER(ctypes.c_double)
_I64 = ctypes.POINTER(ctypes.c_int64)
_lib = None
_checked = False
MIN_N = 4096      # below this numpy's own call is as fast


def _load():
    global _lib, _checked
    if _checked:
        return _lib
    _checked = True
    if os.environ.get("TRI_B200_NUMPY_RNG") or not os.path.exists(_build.HOST_SO_PATH):
        return None
    try:
        L = ctypes.CDLL(_build.HOST_SO_PATH)
        L.trih_mt_rand.argtypes = [_U32, ctypes.POINTER(ctypes.c_int32), _D, ctypes.c_int64,
                                   ctypes.c_int]
        L.trih_mt_randint.argtypes = [_U32, ctypes.POINTER(ctypes.c_int32), ctypes.c_int64,
                                      ctypes.c_uint32, _I64, ctypes.c_int64]
        L.trih_legacy_beta.argtypes = [_U32, ctypes.POINTER(ctypes.c_int32),
                                       ctypes.POINTER(ctypes.c_int32),
                                       ctypes.POINTER(ctypes.c_double), ctypes.c_double,
                                       ctypes.c_double, _D, ctypes.c_int64, ctypes.c_int]
    except (OSError, AttributeError):
        return None
    _lib = L
    if not _self_check():
        _lib = None
    return _lib


def _state():
    st = np.random.get_state()
    if st[0] != "MT19937":
        return None
    return st, np.array(st[1], dtype=np.uint32), ctypes.c_int32(int(st[2]))


def _commit(st, key, pos):
    np.random.set_state((st[0], key, int(pos.value), st[3], st[4]))


def _convert_threads():
    from ._hostpar import N_THREADS
    return max(1, min(int(N_THREADS), 4))


def _advance(n, out):
    """Draw n doubles into `out` (None: only advance the state); False if not possible."""
    s = _state()
    if s is None:
        return False
    st, key, pos = s
    if _lib.trih_mt_rand(key.ctypes.data_as(_U32), ctypes.byref(pos),
                         out.ctypes.data_as(_D) if out is not None else None, n,
                         _convert_threads()) != 0:
        return False
    _commit(st, key, pos)
    return True


def _rand(n, out):
    return out if _advance(n, out) else None


def _self_check():
    """The C continuation against numpy on the same state (state restored afterwards)."""
    saved = np.random.get_state()
    try:
        for seed, n in ((7, 4099), (8, 312), (9, 100_003)):
            np.random.seed(seed)
            np.random.rand(5)
            a, ai, a2 = np.random.rand(n), np.random.randint(0, 2499, 997), np.random.rand(3)
            np.random.seed(seed)
            np.random.rand(5)
            b = _rand(n, np.empty(n))
            bi = randint(0, 2499, 997, _force=True)
            b2 = _rand(3, np.empty(3))
            if b is None or not (np.array_equal(a, b) and np.array_equal(ai, bi)
                                 and np.array_equal(a2, b2)):
                return False
        from scipy.stats import beta
        np.random.seed(10)
        np.random.standard_normal(1)          # a cached gaussian in the state
        a, a2 = beta.rvs(0.867, 3.03, size=70_001), np.random.standard_normal(2)
        np.random.seed(10)
        np.random.standard_normal(1)
        b, b2 = beta_rvs(0.867, 3.03, 70_001, _force=True), np.random.standard_normal(2)
        return bool(np.array_equal(a, b) and np.array_equal(a2, b2))
    finally:
        np.random.set_state(saved)


def _out(n, pinned):
    if pinned:
        from ._bufpool import empty
        return empty(n)
    return np.empty(n)


def rand(n, pinned=False):
    """np.random.rand(n); pinned: on a page-locked buffer (a column that goes to the GPU)."""
    n = int(n)
    if n < MIN_N or _load() is None:
        return np.random.rand(n)
    out = _rand(n, _out(n, pinned))
    return out if out is not None else np.random.rand(n)


def skip(n):
    """Advance the generator as np.random.rand(n) would, without building the array."""
    n = int(n)
    if n < MIN_N or _load() is None or not _advance(n, None):
        np.random.rand(n)


def uniform(low, high, n):
    """np.random.uniform(low, high, n): low + (high - low) * rand (numpy's random_uniform)."""
    n = int(n)
    if n < MIN_N or _load() is None:
        return np.random.uniform(low=low, high=high, size=n)
    out = _rand(n, np.empty(n))
    if out is None:
        return np.random.uniform(low=low, high=high, size=n)
    return float(low) + (float(high) - float(low)) * out


def randint(low, high, n, _force=False):
    n, low, high = int(n), int(low), int(high)
    rng = high - 1 - low
    if (not _force and (n < MIN_N or _load() is None)) or not 0 < rng < 0xFFFFFFFF:
        return np.random.randint(low, high, n)
    s = _state()
    if s is None:
        return np.random.randint(low, high, n)
    st, key, pos = s
    out = np.empty(n, dtype=np.int64)
    if _lib.trih_mt_randint(key.ctypes.data_as(_U32), ctypes.byref(pos), low, rng,
                            out.ctypes.data_as(_I64), n) != 0:
        return np.random.randint(low, high, n)
    _commit(st, key, pos)
    return out


def powerlaw_rvs(a, n):
    """scipy.stats.powerlaw.rvs(a, size=n): the default inverse-CDF sampler of rv_continuous,
    pow(random_state.uniform(size=n), 1/a) (then * scale + loc with scale 1, loc 0)."""
    n = int(n)
    if n < MIN_N or _load() is None:
        from scipy.stats import powerlaw
        return powerlaw.rvs(a, size=n)
    x = rand(n)
    return np.power(x, 1.0 / a, out=x)        # (* scale + loc with 1, 0: the same bits)


def beta_rvs(a, b, n, _force=False, pinned=False):
    """scipy.stats.beta.rvs(a, b, size=n) on numpy's global generator (legacy_beta: two gamma
    deviates by rejection).  The C walker covers a < 1 < b; anything else goes to scipy."""
    n = int(n)
    if (not _force and (n < MIN_N or _load() is None)) or not (0.0 < a < 1.0 < b):
        from scipy.stats import beta
        return beta.rvs(a, b, size=n)
    st = np.random.get_state()
    if st[0] != "MT19937":
        from scipy.stats import beta
        return beta.rvs(a, b, size=n)
    from ._hostpar import N_THREADS
    key, pos = np.array(st[1], dtype=np.uint32), ctypes.c_int32(int(st[2]))
    has_gauss, gauss = ctypes.c_int32(int(st[3])), ctypes.c_double(float(st[4]))
    out = _out(n, pinned)
    rc = _lib.trih_legacy_beta(key.ctypes.data_as(_U32), ctypes.byref(pos),
                               ctypes.byref(has_gauss), ctypes.byref(gauss), float(a), float(b),
                               out.ctypes.data_as(_D), n, int(N_THREADS))
    if rc != 0:     # (state untouched on failure)
        from scipy.stats import beta
        return beta.rvs(a, b, size=n)
    np.random.set_state((st[0], key, int(pos.value), int(has_gauss.value), float(gauss.value)))
    # (rv_continuous.rvs returns vals * scale + loc with scale 1, loc 0: the same bits for
    # these values in (0, 1))
    return out
