"""Host-side timing of the prior draws on this box: numpy / scipy calls against the bulk
continuation (triceratops_b200/_fastrng.py), and the host part of a calc_probs call.
    python scripts/host_rng_bench.py > gpurun_out/host_rng.json"""
import json
import os
import sys
import time

import numpy as np
from scipy.stats import beta, powerlaw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from triceratops_b200 import _fastrng, _hostpar  # noqa: E402

N = 1_000_000


def t(fn, k=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    return (time.perf_counter() - t0) / k * 1e3


assert _fastrng._load() is not None
out = {"threads": _hostpar.N_THREADS, "cores": os.cpu_count(), "ms_per_1e6": {
    "np.random.rand": t(lambda: np.random.rand(N)), "fast rand": t(lambda: _fastrng.rand(N)),
    "fast skip": t(lambda: _fastrng.skip(N)),
    "np.random.randint": t(lambda: np.random.randint(0, 2499, N)),
    "fast randint": t(lambda: _fastrng.randint(0, 2499, N)),
    "scipy powerlaw.rvs": t(lambda: powerlaw.rvs(0.2, size=N)),
    "fast powerlaw": t(lambda: _fastrng.powerlaw_rvs(0.2, N)),
    "scipy beta.rvs": t(lambda: beta.rvs(0.867, 3.03, size=N)),
    "fast beta": t(lambda: _fastrng.beta_rvs(0.867, 3.03, N))}}
import _workloads  # noqa: E402
lc = _workloads.lightcurve(2)
for mode in ("fast", "numpy"):
    if mode == "numpy":
        _fastrng._lib = None
    ts = [_workloads.record_calls(2, N, 2026, lc)[1] for _ in range(3)]
    out["calc_probs_host_s_" + mode] = ts
print(json.dumps(out))
