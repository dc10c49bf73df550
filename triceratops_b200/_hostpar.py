"""Host-side parallelism for the prior-draw preparation (never for the light-curve path).

The draws themselves come from numpy's global RNG and stay sequential (bit-identical streams);
what is spread over cores are the deterministic element-wise transforms applied to them, in two
ways that both leave every element's value unchanged:

  * `splev`: FITPACK spline evaluation through csrc/host_prep.c (same recurrence, OpenMP) instead
    of scipy's GIL-bound single-threaded call;
  * `pmap`: an element-wise numpy function applied to contiguous chunks in a thread pool (numpy
    releases the GIL inside ufunc loops).
"""
import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _build

_D = ctypes.POINTER(ctypes.c_double)
MIN_CHUNK = 1 << 16


def _n_threads():
    v = os.environ.get("TRI_B200_HOST_THREADS")
    if v:
        return max(1, int(v))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        cores = os.cpu_count() or 1
    # one process per GPU: the ranks of a node share its cores.  With numpy's draws rank 0 makes
    # them for every rank (the others wait in the scatter), so it takes what the others leave
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    if ranks > 1 and os.environ.get("LOCAL_RANK", "0") == "0":
        return max(1, min(cores - (ranks - 1), 32))
    return max(1, min(cores // ranks, 32))


def _tune_malloc():
    """TRI_B200_MALLOC_TUNE=1: keep large host arrays in glibc's heap instead of fresh mmap
    regions, so that the 8 MB draw columns of consecutive scenarios reuse already-faulted pages
    (a first-touch page fault per 4 KB is ~40 % of a bulk `rand(1e6)`).  Off by default: it is a
    process-wide allocator setting and keeps freed memory in the process."""
    if os.environ.get("TRI_B200_MALLOC_TUNE") != "1":
        return False
    try:
        libc = ctypes.CDLL("libc.so.6")
        M_TRIM_THRESHOLD, M_MMAP_THRESHOLD = -1, -3
        return bool(libc.mallopt(M_MMAP_THRESHOLD, 1 << 30)
                    and libc.mallopt(M_TRIM_THRESHOLD, (1 << 31) - 1))
    except OSError:
        return False


MALLOC_TUNED = _tune_malloc()
N_THREADS = _n_threads()
_pool = ThreadPoolExecutor(N_THREADS) if N_THREADS > 1 else None
_lib = None
_lib_tried = False


def _host_lib():
    """csrc/libtriceratops_host.so, or None: scipy then evaluates the same spline itself (the
    values are identical either way; this is host-side preparation, not the GPU path)."""
    global _lib, _lib_tried
    if not _lib_tried:
        _lib_tried = True
        if os.path.exists(_build.HOST_SO_PATH):
            L = ctypes.CDLL(_build.HOST_SO_PATH)
            L.trih_splev.argtypes = [_D, ctypes.c_int, _D, ctypes.c_int, _D, _D, ctypes.c_int64,
                                     ctypes.c_int]
            _lib = L
    return _lib


def splev(spline, x):
    """spline(x) for a scipy InterpolatedUnivariateSpline, bit-identical, multi-threaded."""
    x = np.asarray(x, dtype=np.float64)
    L = _host_lib()
    if L is None or x.ndim != 1 or x.size < 256:
        return spline(x)
    t, c, k = spline._eval_args
    t = np.ascontiguousarray(t, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    x = np.ascontiguousarray(x)
    y = np.empty_like(x)
    rc = L.trih_splev(t.ctypes.data_as(_D), t.size, c.ctypes.data_as(_D), int(k),
                      x.ctypes.data_as(_D), y.ctypes.data_as(_D), x.size, N_THREADS)
    if rc != 0:
        return spline(x)
    return y


def slices(n):
    """Contiguous chunk boundaries for n elements (one chunk when small or single-threaded)."""
    if _pool is None or n < 2 * MIN_CHUNK:
        return [slice(0, n)]
    parts = min(N_THREADS, max(1, n // MIN_CHUNK))
    edges = [(i * n) // parts for i in range(parts + 1)]
    return [slice(edges[i], edges[i + 1]) for i in range(parts)]


def pmap(fn, n, *arrays):
    """fn(*[a[chunk] for a in arrays]) over chunks in parallel; arrays that are not length-n
    ndarrays are passed whole.  Returns the list of per-chunk results (in order)."""
    sl = slices(n)

    def pick(a, s):
        return a[s] if isinstance(a, np.ndarray) and a.ndim >= 1 and a.shape[0] == n else a

    if len(sl) == 1:
        return [fn(*arrays)]
    futs = [_pool.submit(fn, *[pick(a, s) for a in arrays]) for s in sl]
    return [f.result() for f in futs]


def pmap_concat(fn, n, *arrays):
    """pmap for functions returning one array (or a tuple of arrays) per chunk."""
    res = pmap(fn, n, *arrays)
    if len(res) == 1:
        return res[0]
    if isinstance(res[0], tuple):
        return tuple(np.concatenate([r[i] for r in res]) for i in range(len(res[0])))
    return np.concatenate(res)
