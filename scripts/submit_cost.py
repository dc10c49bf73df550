"""Host time spent inside tri_submit_tp for pageable and for page-locked caller columns (the
call must not wait for the GPU): python scripts/submit_cost.py (needs a GPU)."""
import sys, os, time, ctypes, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import draw_tp_columns, TOI465, GOLD
from triceratops_b200 import _cabi
from triceratops_b200._cabi import tri_result, tri_tp_args
from triceratops_b200.engine import get_engine
eng = get_engine()
lc = np.loadtxt(os.path.join(GOLD, "TOI465_01_lightcurve.csv"), delimiter=",")
eng.set_lightcurve(lc[:, 0].copy(), lc[:, 1].copy(), float(np.mean(lc[:, 2])), 0.00139, 20)
N = 1_000_000
cols = draw_tp_columns(N, 1)
def make_args(pinned):
    a = tri_tp_args(); a.N = N; keep = []
    for name, val in cols.items():
        arr = np.ascontiguousarray(np.broadcast_to(np.asarray(val, float), (N,)))
        if pinned:
            t = torch.from_numpy(arr).pin_memory(); keep.append(t); arr = t.numpy()
        keep.append(arr)
        setattr(a, name, _cabi.tri_col(arr.ctypes.data, 1))
    return a, keep
for pinned in (False, True, False, True):
    a, keep = make_args(pinned)
    for rep in range(4):
        r = (tri_result * 1)(); tk = ctypes.c_int64()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _cabi.check(eng.lib.tri_submit_tp(ctypes.byref(a), r, ctypes.byref(tk)))
        t1 = time.perf_counter()
        _cabi.check(eng.lib.tri_wait(tk, r))
        t2 = time.perf_counter()
        print("pinned" if pinned else "pageable", "submit %.2f ms  wait %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), r[0].lnZ)
print("host threads env", os.environ.get("TRI_B200_HOST_THREADS"), "cpus", os.cpu_count())
