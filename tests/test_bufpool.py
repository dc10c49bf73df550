"""Recycled page-locked buffers (_bufpool): a buffer is reused only after the array on it and
every view derived from that array are gone; without CUDA the pool is plain numpy."""
import gc

import numpy as np
import pytest

from triceratops_b200 import _bufpool


def test_without_cuda_or_for_small_arrays_it_is_numpy():
    assert _bufpool.empty(1000).flags.owndata
    import torch
    if not torch.cuda.is_available():
        a = _bufpool.empty(400_000)
        assert a.flags.owndata and a.shape == (400_000,)


def test_lease_follows_the_last_view(monkeypatch):
    # (a stand-in for the pinned allocator: the lease logic does not depend on where the bytes live)
    monkeypatch.setattr(_bufpool, "_pinned_owner", lambda nbytes: np.empty(nbytes, dtype=np.uint8))
    monkeypatch.setattr(_bufpool, "_free", {})
    monkeypatch.setattr(_bufpool, "_calls", [2])
    n = 300_000
    a = _bufpool.empty(n)
    assert a.shape == (n,) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable
    a[:] = 7.0
    addr = a.ctypes.data
    v = a[1000:2000].reshape(10, 100).T          # a view of a view of a view
    del a
    gc.collect()
    b = _bufpool.empty(n)                        # the first buffer is still leased by v
    assert b.ctypes.data != addr and float(v[3, 4]) == 7.0
    del v
    gc.collect()
    c = _bufpool.empty(n)                        # now it comes back
    assert c.ctypes.data == addr
    d = _bufpool.empty(n, np.int64)
    assert d.dtype == np.int64 and d.ctypes.data not in (b.ctypes.data, c.ctypes.data)


@pytest.mark.gpu
def test_pooled_columns_are_page_locked_and_skip_the_staging_copy(gpu_engine):
    import torch
    assert _bufpool.empty(500_000).flags.owndata or _bufpool._calls[0] >= 2    # off before a 2nd call
    _bufpool.note_call()
    _bufpool.note_call()
    a = _bufpool.empty(500_000)
    assert not a.flags.owndata
    t = torch.from_numpy(a)
    assert t.is_pinned()
