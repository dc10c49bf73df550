"""tri_submit_* / tri_wait: evaluations queued back to back give the answers of the synchronous
calls, in any wait order, and the ring of TRI_MAX_INFLIGHT slots is enforced."""
import ctypes

import numpy as np
import pytest

from conftest import draw_tp_columns, draw_eb_columns

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert a.lnZ == b.lnZ and a.n_pass == b.n_pass and a.n_evaluated == b.n_evaluated
    np.testing.assert_array_equal(a.top_idx, b.top_idx)
    np.testing.assert_array_equal(a.top_lnL, b.top_lnL)
    if a.lnL is not None and b.lnL is not None:
        np.testing.assert_array_equal(a.lnL, b.lnL)


def test_pipelined_calls_match_synchronous_ones(gpu_engine, toi465_lc):
    eng = gpu_engine
    t, f, s = toi465_lc
    eng.set_lightcurve(t, f, s, 0.00139, 20)
    N = 30000
    tp = [draw_tp_columns(N, seed) for seed in (1, 2, 3)]
    eb = [draw_eb_columns(N, seed) for seed in (4, 5)]
    sync_tp = [eng.eval_tp(N, **c, want_lnL=True, n_best=100) for c in tp]
    sync_eb = [eng.eval_eb(N, **c, want_lnL=True, n_best=100) for c in eb]
    # queue everything (5 calls > 4 slots: the engine waits for the oldest by itself) ...
    pend = [eng.submit_tp(N, **tp[0], want_lnL=True, n_best=100),
            eng.submit_eb(N, **eb[0], want_lnL=True, n_best=100),
            eng.submit_tp(N, **tp[1], want_lnL=True, n_best=100),
            eng.submit_eb(N, **eb[1], want_lnL=True, n_best=100),
            eng.submit_tp(N, **tp[2], want_lnL=True, n_best=100)]
    # ... and read the results out of order
    _same(pend[4].result(), sync_tp[2])
    _same(pend[0].result(), sync_tp[0])
    for got, want in zip(pend[3].result(), sync_eb[1]):
        _same(got, want)
    _same(pend[2].result(), sync_tp[1])
    for got, want in zip(pend[1].result(), sync_eb[0]):
        _same(got, want)
    assert not eng._inflight


def test_slot_ring_is_enforced_by_the_library(gpu_engine, toi465_lc):
    from triceratops_b200 import _cabi
    from triceratops_b200._cabi import tri_result, tri_tp_args
    eng = gpu_engine
    t, f, s = toi465_lc
    eng.set_lightcurve(t, f, s, 0.00139, 20)
    N = 2000
    cols = draw_tp_columns(N, 7)
    a = tri_tp_args()
    a.N = N
    keep = []
    for name, val in cols.items():
        arr = np.ascontiguousarray(np.broadcast_to(np.asarray(val, float), (N,)))
        keep.append(arr)
        setattr(a, name, _cabi.tri_col(arr.ctypes.data, 1))
    tickets, results = [], []
    for _ in range(_cabi.TRI_MAX_INFLIGHT):
        r = (tri_result * 1)()
        tk = ctypes.c_int64()
        _cabi.check(eng.lib.tri_submit_tp(ctypes.byref(a), r, ctypes.byref(tk)))
        tickets.append(tk.value)
        results.append(r)
    assert len(set(tickets)) == len(tickets)
    r = (tri_result * 1)()
    tk = ctypes.c_int64()
    assert eng.lib.tri_submit_tp(ctypes.byref(a), r, ctypes.byref(tk)) == _cabi.TRI_ESTATE
    for tkt, res in zip(tickets, results):
        _cabi.check(eng.lib.tri_wait(ctypes.c_int64(tkt), res))
    assert len({res[0].lnZ for res in results}) == 1          # same draws, same answer
    # a ticket can be waited for once
    assert eng.lib.tri_wait(ctypes.c_int64(tickets[0]), results[0]) == _cabi.TRI_EINVAL
    # and the ring is free again
    _cabi.check(eng.lib.tri_submit_tp(ctypes.byref(a), r, ctypes.byref(tk)))
    _cabi.check(eng.lib.tri_wait(tk, r))
    assert r[0].lnZ == results[0][0].lnZ


def test_new_light_curve_waits_for_calls_in_flight(gpu_engine, toi465_lc):
    eng = gpu_engine
    t, f, s = toi465_lc
    N = 20000
    cols = draw_tp_columns(N, 9)
    eng.set_lightcurve(t, f, s, 0.00139, 20)
    want = eng.eval_tp(N, **cols, want_lnL=True)
    p = eng.submit_tp(N, **cols, want_lnL=True)
    eng.set_lightcurve(t, f + 1e-3, s, 0.00139, 20)   # must not disturb the call in flight
    other = eng.eval_tp(N, **cols, want_lnL=True)
    got = p.result()
    np.testing.assert_array_equal(got.lnL, want.lnL)
    assert other.lnZ != want.lnZ
