"""Where the wall time of a host-sampler calc_probs goes: per scenario, the time its draws hold
numpy's generator (the sequential chain) and the time of its deterministic part."""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import _workloads  # noqa: E402
from triceratops_b200 import _dispatch, _fastrng  # noqa: E402
from triceratops_b200.engine import get_engine  # noqa: E402

get_engine(0)
lc = _workloads.lightcurve(2)
tgt = _workloads.make_target(2)
for _ in range(2):
    _workloads.run_calc_probs(tgt, 2, lc, 1_000_000, 2026)
events = []
lock = threading.Lock()
orig_run = _dispatch.ScenarioChain.run
orig_done = _dispatch.rng_done
rng_time = {}


def wrap_rng(name):
    f = getattr(_fastrng, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        with lock:
            rng_time[name] = rng_time.get(name, 0.0) + time.perf_counter() - t0
        return r
    setattr(_fastrng, name, g)


for n in ("rand", "skip", "randint", "powerlaw_rvs", "beta_rvs", "uniform"):
    wrap_rng(n)


def run(self, fn):
    def traced():
        t0 = time.perf_counter()
        tl = threading.local()
        out = fn()
        with lock:
            events.append(("scenario", getattr(fn, "func", fn).__name__, t0, time.perf_counter()))
        return out
    return orig_run(self, traced)


def done():
    with lock:
        events.append(("rng_done", threading.get_ident(), time.perf_counter(), 0))
    orig_done()


_dispatch.ScenarioChain.run = run
_dispatch.rng_done = done
import triceratops_b200.marginal_likelihoods as ml  # noqa: E402
ml._dispatch.rng_done = done
t0 = time.perf_counter()
_workloads.run_calc_probs(tgt, 2, lc, 1_000_000, 2026)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
sc = sorted([e for e in events if e[0] == "scenario"], key=lambda e: e[2])
dn = sorted([e[2] for e in events if e[0] == "rng_done"])
print(json.dumps({"wall_s": wall, "rng_seconds_by_call": rng_time, "rng_total": sum(rng_time.values()),
                  "scenarios": [(n, round(a - t0, 3), round(b - t0, 3)) for _, n, a, b in sc],
                  "rng_done_at": [round(x - t0, 3) for x in dn]}))
