"""Recycled page-locked host buffers for the columns that go to the GPU.

The engine uploads a column straight from the caller's buffer when that buffer is page-locked
and through a pinned staging copy otherwise (`via_pinned`, csrc/tri_cabi.cu) -- ~100 MB of
memcpy per scenario.  Page-locking is far too slow to do per array, so `empty()` hands out
arrays on pinned buffers that return to a free list when the last view of them dies: a warm
process neither allocates nor stages.

The array returned is a view (`owndata` False) of a holder object that keeps the buffer leased
for as long as the array or any view derived from it is alive.  TRI_B200_PINNED_POOL_MB caps the
pinned memory the pool may own (default 3072; 0 disables it); beyond the cap, and wherever CUDA
is not available, `empty()` is `numpy.empty`.  Page-locking ~1.3 GB costs ~0.5 s once, so the
pool only switches on with a process's SECOND `calc_probs` (`note_call()`): a script that vets
one target is not slowed down, a sweep gains from its third call on.
"""
import collections
import ctypes
import os

import numpy as np

MIN_BYTES = 1 << 20
_GRAIN = 1 << 16
try:
    _CAP = max(0, int(os.environ.get("TRI_B200_PINNED_POOL_MB", "3072"))) << 20
except ValueError:
    _CAP = 3072 << 20
_free = {}               # nbytes -> deque of owner arrays (deque.append / pop are atomic)
_holders = {}            # nbytes -> ctypes holder type
_owned = [0]             # pinned bytes allocated by the pool (leased or free)
_usable = [None]         # None: not tried; False: no pinned memory here
stats = {"new": 0, "reused": 0, "plain": 0}
_calls = [0]


def note_call():
    """calc_probs tells the pool that another public call starts (see the module docstring)."""
    _calls[0] += 1


def _holder_type(nbytes):
    T = _holders.get(nbytes)
    if T is None:
        base = ctypes.c_uint8 * nbytes

        class Holder(base):
            """Lease on one pooled buffer; returns it when the last array on it is gone."""
            _owner = None

            def __del__(self):
                try:
                    owner, self._owner = self._owner, None
                    if owner is not None:
                        _free.setdefault(owner.nbytes, collections.deque()).append(owner)
                except Exception:      # interpreter shutdown: the module's globals are gone
                    pass

        T = _holders.setdefault(nbytes, Holder)
    return T


def _pinned_owner(nbytes):
    if _usable[0] is False or _owned[0] + nbytes > _CAP:
        return None
    try:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device")
        owner = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True).numpy()
    except Exception:
        _usable[0] = False
        return None
    _usable[0] = True
    _owned[0] += nbytes
    stats["new"] += 1
    return owner


def empty(n, dtype=np.float64):
    """Uninitialised 1-D array of n elements, like np.empty(n, dtype), page-locked if possible."""
    dtype = np.dtype(dtype)
    want = int(n) * dtype.itemsize
    if want < MIN_BYTES or _CAP == 0 or _calls[0] < 2:
        return np.empty(int(n), dtype=dtype)
    nbytes = -(-want // _GRAIN) * _GRAIN
    owner = None
    q = _free.get(nbytes)
    if q is not None:
        try:
            owner = q.pop()
            stats["reused"] += 1
        except IndexError:
            owner = None
    if owner is None:
        owner = _pinned_owner(nbytes)
    if owner is None:
        stats["plain"] += 1
        return np.empty(int(n), dtype=dtype)
    holder = _holder_type(nbytes).from_buffer(owner)
    holder._owner = owner
    return np.frombuffer(holder, dtype=dtype, count=int(n))


def release_all():
    """Give the free buffers back (leased ones follow when their arrays die and are dropped by
    the next release_all)."""
    for q in list(_free.values()):
        while True:
            try:
                _owned[0] -= q.pop().nbytes
            except IndexError:
                break
