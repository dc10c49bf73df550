"""Multi-rank path on CPU: world_size 2 over gloo.  Single lnZ_* calls: every rank makes the
same host draws, evaluates its own contiguous slice and one all-gather merges the (max,
scaled-sum) records and the best-draw candidates.  calc_probs: rank 0 draws and scatters every
engine call's columns, ONE all-gather carries the records of all rows.  Either way the result
must equal the single-process run."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np

from conftest import RESULT_KEYS, ROOT, TOI465


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_equal_one(tmp_path, oracle_engine, toi465_lc):
    out = str(tmp_path / "res.pkl")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2",
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests",
                                                                    "_dist_worker.py"), out],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    res = [pickle.load(open(out + ".%d" % r, "rb")) for r in range(2)]
    assert res[0]["shard"] == (0, 1000) and res[1]["shard"] == (1000, 2001)

    import triceratops_b200.marginal_likelihoods as ml
    t, f, s = toi465_lc
    N = 2001
    np.random.seed(123)
    tp = ml.lnZ_TTP(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0, N, True)
    np.random.seed(124)
    eb = ml.lnZ_TEB(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0, N, True)
    np.random.seed(125)
    ptp = ml.lnZ_PTP(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0,
                     TOI465["plx"], None, "TESS", N, True)

    def same(a, b):
        assert abs(a["lnZ"] - b["lnZ"]) < 1e-9, (a["lnZ"], b["lnZ"])
        assert (a.n_pass, a.n_evaluated) == (b.n_pass, b.n_evaluated)
        # rows beyond the evaluated draws carry zero weight and arbitrary order
        k_rows = min(100, b.n_evaluated)
        for k in RESULT_KEYS:
            np.testing.assert_allclose(a[k][:k_rows], b[k][:k_rows], rtol=1e-12, err_msg=k)

    from conftest import calc_probs_small
    one = calc_probs_small(t, f, s, full=True)
    lnZ_cp = one.lnZ
    assert np.isfinite(lnZ_cp).sum() >= 3 and one.collectives is None
    for r in res:
        np.testing.assert_allclose(r["lnZ_cp"], lnZ_cp, rtol=0, atol=1e-9)
        # the whole table reaches every rank, from one exchange of records (+ one scatter of
        # draw columns per engine call: TTP, TEB, DTP)
        np.testing.assert_allclose(r["cp"]["prob"], one.probs.prob.values, rtol=0, atol=1e-9)
        np.testing.assert_allclose(r["cp"]["R_p"], one.probs.R_p.values, rtol=1e-12)
        np.testing.assert_allclose(r["cp"]["inc"], one.probs.inc.values, rtol=1e-12)
        np.testing.assert_allclose(r["cp"]["u1"], one.u1, rtol=0, atol=0)
        assert abs(r["cp"]["FPP"] - one.FPP) < 1e-9
        assert r["cp"]["collectives"] == {"record_exchanges": 1, "scatters": 3}

    for r in res:                       # both ranks hold the combined answer
        same(r["tp"], tp)
        same(r["ptp"], ptp)
        same(r["eb"][0], eb[0])
        same(r["eb"][1], eb[1])
