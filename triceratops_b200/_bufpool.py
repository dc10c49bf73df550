"""Recycled host buffers for the per-draw arrays of the prior-draw preparation.

A calc_probs at N = 1e6 creates ~250 arrays of 8 MB (deviates, transformed columns) and drops
them at the end of the call.  glibc serves each from a fresh mmap and returns it with munmap:
every array is page-faulted in again (the kernel zeroes 2 GB per call) and every munmap stalls
the other threads of the process (mmap lock, TLB shootdown) -- with the generator and the
preparation in C that churn was the largest single item left in the call (-12 % wall with
glibc told not to mmap).  `empty()` hands out arrays on buffers that go back to a free list when
the last view of them dies, so a warm process allocates nothing.

The array returned is a view (`owndata` False) of a holder object that keeps the buffer leased
for as long as the array or any view derived from it is alive.  TRI_B200_POOL_MB caps the memory
kept for reuse (default 4096; 0 disables the pool).
"""
import collections
import ctypes
import os

import numpy as np

MIN_BYTES = 1 << 20
_GRAIN = 1 << 16
try:
    _CAP = max(0, int(os.environ.get("TRI_B200_POOL_MB", "4096"))) << 20
except ValueError:
    _CAP = 4096 << 20
_free = {}               # nbytes -> deque of owner arrays (deque.append / pop are atomic)
_holders = {}            # nbytes -> ctypes holder type
_retained = [0]
stats = {"new": 0, "reused": 0, "dropped": 0}


def _holder_type(nbytes):
    T = _holders.get(nbytes)
    if T is None:
        base = ctypes.c_uint8 * nbytes

        class Holder(base):
            """Lease on one pooled buffer; returns it when the last array on it is gone."""
            _owner = None

            def __del__(self):
                owner, self._owner = self._owner, None
                if owner is None:
                    return
                if _retained[0] + owner.nbytes <= _CAP:
                    _retained[0] += owner.nbytes
                    _free.setdefault(owner.nbytes, collections.deque()).append(owner)
                else:
                    stats["dropped"] += 1

        T = _holders.setdefault(nbytes, Holder)
    return T


def empty(n, dtype=np.float64):
    """Uninitialised 1-D array of n elements, like np.empty(n, dtype)."""
    dtype = np.dtype(dtype)
    want = int(n) * dtype.itemsize
    if want < MIN_BYTES or _CAP == 0:
        return np.empty(int(n), dtype=dtype)
    nbytes = -(-want // _GRAIN) * _GRAIN
    owner = None
    q = _free.get(nbytes)
    if q is not None:
        try:
            owner = q.pop()
            _retained[0] -= nbytes
            stats["reused"] += 1
        except IndexError:
            owner = None
    if owner is None:
        owner = np.empty(nbytes, dtype=np.uint8)
        stats["new"] += 1
    holder = _holder_type(nbytes).from_buffer(owner)
    holder._owner = owner
    return np.frombuffer(holder, dtype=dtype, count=int(n))


def release_all():
    """Forget the buffers kept for reuse (their memory goes back to the allocator)."""
    for q in list(_free.values()):
        while True:
            try:
                q.pop()
            except IndexError:
                break
    _retained[0] = 0
