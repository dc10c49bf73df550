"""What slows the generator helpers down inside the calc_probs pipeline?  Times
_fastrng.beta_rvs / rand alone and beside synthetic background load of three kinds in other
threads of the same process: scalar arithmetic (cores), streaming writes (memory bandwidth),
and the real preparation blocks.  One JSON line."""
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from triceratops_b200 import _fastrng, _hostpar  # noqa: E402

N = 1_000_000
np.random.seed(1)


def timed(fn, k=8):
    fn()
    ts = []
    for _ in range(k):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return round(float(np.median(ts)) * 1e3, 2)


def measure():
    return {"beta_ms": timed(lambda: _fastrng.beta_rvs(0.867, 3.03, N)),
            "rand_ms": timed(lambda: _fastrng.rand(N), 20)}


stop = threading.Event()


def spin_load():                      # numpy ufunc on a cache-resident array: cores only
    a = np.random.rand(4096)
    while not stop.is_set():
        for _ in range(200):
            np.sqrt(a, out=a)
            np.add(a, 1.0, out=a)


def stream_load():                    # 64 MB streaming copies: memory bandwidth
    a = np.zeros(8_000_000)
    b = np.zeros(8_000_000)
    while not stop.is_set():
        np.copyto(b, a)


def block_load():                     # the real thing: a scenario's preparation block
    import triceratops_b200.marginal_likelihoods as ml
    from triceratops_b200 import _blocks
    _blocks.available()
    rng = np.random.default_rng(3)
    x = [rng.random(N) for _ in range(5)]
    while not stop.is_set():
        y = [v.copy() for v in x]
        _blocks.run("PEB", N, M_s=0.93, R_s=0.95, Teff=5400.0, c_comp=y[0], x_inc=y[1], x_q=y[2],
                    x_e=y[3], x_w=y[4], P_mean=4.2, plx=8.1, bound_kind="EB", _force=True)


out = {"host_threads": _hostpar.N_THREADS, "alone": measure()}
for name, fn, n in (("spin x8", spin_load, 8), ("spin x16", spin_load, 16),
                    ("stream x4", stream_load, 4), ("stream x8", stream_load, 8),
                    ("one preparation block in a loop", block_load, 1)):
    stop.clear()
    th = [threading.Thread(target=fn, daemon=True) for _ in range(n)]
    for t in th:
        t.start()
    time.sleep(0.3)
    out[name] = measure()
    stop.set()
    for t in th:
        t.join()
print(json.dumps(out))
