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
import threading
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
_tl = threading.local()      # .inline: this thread is already one of a block's workers
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
            L.trih_take_f64.argtypes = [_D, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), _D,
                                        ctypes.c_int64, ctypes.c_int]
            _lib = L
    return _lib


def take(table, idx):
    """table[idx] for a 1-D float64 table and int64 indices: the same values as numpy's fancy
    indexing, gathered without the GIL over the host threads (numpy holds the GIL for the ~5 ms
    a 1e6-element gather takes, which stalls the thread drawing the next scenario's priors)."""
    L = _host_lib()
    if (L is None or not isinstance(table, np.ndarray) or not isinstance(idx, np.ndarray)
            or table.dtype != np.float64 or idx.dtype != np.int64 or table.ndim != 1
            or idx.ndim != 1 or idx.size < 4096
            or not table.flags.c_contiguous or not idx.flags.c_contiguous):
        return table[idx]
    out = np.empty(idx.size)
    rc = L.trih_take_f64(table.ctypes.data_as(_D), table.size,
                         idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                         out.ctypes.data_as(_D), idx.size, _c_threads(min(N_THREADS, 8)))
    if rc != 0:
        return table[idx]          # (raises numpy's own IndexError)
    return out


def _inline():
    return getattr(_tl, "inline", False)


def _c_threads(n):
    """OpenMP threads for a C helper: one inside a block worker (the block is the parallelism)."""
    return 1 if _inline() else n


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
                      x.ctypes.data_as(_D), y.ctypes.data_as(_D), x.size, _c_threads(N_THREADS))
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

    if len(sl) == 1 or _inline():
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


def block_chunks(n):
    """Chunk boundaries of pmap_block (TRI_B200_BLOCK_CHUNKS overrides the count)."""
    if _pool is None or n < 2 * MIN_CHUNK:
        return [slice(0, n)]
    want = os.environ.get("TRI_B200_BLOCK_CHUNKS")
    # one chunk per thread, and more of them when that would push a chunk's temporaries
    # (8 bytes x ~10 live arrays per element) out of a core's cache
    parts = int(want) if want else max(N_THREADS, -(-n // (2 * MIN_CHUNK)))
    parts = max(1, min(parts, n // (MIN_CHUNK // 2)))
    edges = [(i * n) // parts for i in range(parts + 1)]
    return [slice(edges[i], edges[i + 1]) for i in range(parts)]


def pmap_block(fn, n, *arrays):
    """A scenario's whole element-wise preparation, chunk by chunk over the host threads.

    fn(*chunks) -> tuple of per-draw arrays (entries may be None); every length-n ndarray in
    `arrays` is passed as its chunk, anything else whole.  Each worker runs the complete chain
    of numpy passes on a chunk that stays in its core's cache (temporaries of ~0.5 MB come from
    the heap instead of freshly mapped pages) and writes its results into the shared outputs;
    helpers called inside (pmap, splev, take) run inline.  Element-wise operations only: the
    values do not depend on the chunking.
    """
    sl = block_chunks(n)
    if len(sl) == 1 or _inline():
        return fn(*arrays)

    def pick(a, s):
        return a[s] if isinstance(a, np.ndarray) and a.ndim >= 1 and a.shape[0] == n else a

    outs, lock = [], threading.Lock()

    def work(s):
        _tl.inline = True
        try:
            res = fn(*[pick(a, s) for a in arrays])
        finally:
            _tl.inline = False
        with lock:
            if not outs:
                outs.extend(None if r is None else np.empty((n,) + np.shape(r)[1:],
                                                            dtype=np.asarray(r).dtype)
                            for r in res)
        for o, r in zip(outs, res):
            if o is not None:
                o[s] = r
        return len(res)

    futs = [_pool.submit(work, s) for s in sl]
    for f in futs:
        f.result()
    return tuple(outs)
