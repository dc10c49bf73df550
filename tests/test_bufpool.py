"""Recycled host buffers (_bufpool): an array's memory is reused only after the array and every
view derived from it are gone."""
import gc

import numpy as np

from triceratops_b200 import _bufpool


def test_lease_follows_the_last_view():
    _bufpool.release_all()
    n = 300_000
    a = _bufpool.empty(n)
    assert a.shape == (n,) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable
    a[:] = 7.0
    addr = a.ctypes.data
    v = a[1000:2000].reshape(10, 100).T          # a view of a view of a view
    del a
    gc.collect()
    b = _bufpool.empty(n)                        # the first buffer is still leased by v
    assert b.ctypes.data != addr
    assert float(v[3, 4]) == 7.0
    del v
    gc.collect()
    c = _bufpool.empty(n)                        # now it comes back
    assert c.ctypes.data == addr
    d = _bufpool.empty(n, np.int64)              # same size class, another dtype
    assert d.dtype == np.int64 and d.ctypes.data not in (b.ctypes.data, c.ctypes.data)


def test_small_arrays_and_disabled_pool_are_plain_numpy(monkeypatch):
    assert _bufpool.empty(1000).flags.owndata
    monkeypatch.setattr(_bufpool, "_CAP", 0)
    assert _bufpool.empty(400_000).flags.owndata


def test_cap_bounds_what_is_kept(monkeypatch):
    _bufpool.release_all()
    monkeypatch.setattr(_bufpool, "_CAP", 5 << 20)
    arrs = [_bufpool.empty(300_000) for _ in range(4)]     # 2.4 MB each
    del arrs
    gc.collect()
    assert _bufpool._retained[0] <= 5 << 20
    assert sum(len(q) for q in _bufpool._free.values()) == 2
    _bufpool.release_all()
    assert _bufpool._retained[0] == 0
