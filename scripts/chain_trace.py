"""Where the wall time of a host-sampler calc_probs goes.

numpy's global generator makes the scenarios' prior draws strictly sequential (the "chain");
everything else runs beside it.  This wraps the chain's pieces (no product code is changed) and
prints, per scenario: when its thread got the generator, when it released it (`rng_done`), how
much of that span was spent inside the generator calls and what else the thread did meanwhile;
then the tail after the last draw.

    python scripts/chain_trace.py [--draws N] [--config K] [--json out.json]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402,F401
import torch  # noqa: E402
import _workloads  # noqa: E402
import triceratops_b200.marginal_likelihoods as ml  # noqa: E402
from triceratops_b200 import _dispatch, _fastrng, _hostpar, engine as engine_mod  # noqa: E402
from triceratops_b200.engine import get_engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--draws", type=int, default=1_000_000)
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--json", default=None)
args = ap.parse_args()

get_engine(0)
lc = _workloads.lightcurve(args.config)
tgt = _workloads.make_target(args.config)
for _ in range(2):
    _workloads.run_calc_probs(tgt, args.config, lc, args.draws, 2026)

spans = []          # (thread, label, t0, t1)
lock = threading.Lock()
tracing = False


def span(label, t0, t1):
    if tracing:
        with lock:
            spans.append((threading.get_ident(), label, t0, t1))


def wrap(obj, name, label=None):
    f = getattr(obj, name)
    lab = label or name

    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            span(lab, t0, time.perf_counter())
    g.__name__ = getattr(f, "__name__", name)
    setattr(obj, name, g)


RNG = ("rand", "skip", "randint", "powerlaw_rvs", "beta_rvs", "uniform")
for n in RNG:
    wrap(_fastrng, n, "rng:" + n)
for n in ("splev", "pmap", "take"):
    wrap(_hostpar, n, "hostpar:" + n)
for n in ("stellar_relations", "flux_relation", "_background_prior", "_Background", "_periods",
          "_run_tp", "_run_eb", "_fluxratio", "sample_inc", "sample_q", "sample_w", "sample_rp",
          "lnprior_bound_TP", "lnprior_bound_EB", "lnprior_background"):
    if hasattr(ml, n):
        wrap(ml, n, "ml:" + n)
wrap(engine_mod.Engine, "submit_tp", "eng:submit_tp")
wrap(engine_mod.Engine, "submit_eb", "eng:submit_eb")
wrap(engine_mod.Pending, "result", "eng:result")

orig_block = _hostpar.pmap_block


def traced_block(fn, n, *arrays):
    t0 = time.perf_counter()

    def chunk(*a):
        c0 = time.perf_counter()
        try:
            return fn(*a)
        finally:
            span("chunk", c0, time.perf_counter())
    try:
        return orig_block(chunk, n, *arrays)
    finally:
        span("block", t0, time.perf_counter())


_hostpar.pmap_block = traced_block
ml._hostpar.pmap_block = traced_block

orig_run = _dispatch.ScenarioChain.run
orig_done = _dispatch.rng_done


def run(self, fn):
    def traced():
        t0 = time.perf_counter()          # the previous scenario released the generator
        try:
            return fn()
        finally:
            span("scenario:" + getattr(getattr(fn, "func", fn), "__name__", "?"), t0,
                 time.perf_counter())
    return orig_run(self, traced)


def done():
    t = time.perf_counter()
    span("rng_done", t, t)
    orig_done()


_dispatch.ScenarioChain.run = run
_dispatch.rng_done = done
ml._dispatch.rng_done = done

tracing = True
t_begin = time.perf_counter()
_workloads.run_calc_probs(tgt, args.config, lc, args.draws, 2026)
torch.cuda.synchronize()
wall = time.perf_counter() - t_begin
tracing = False

by_thread = {}
for th, lab, a, b in spans:
    by_thread.setdefault(th, []).append((lab, a - t_begin, b - t_begin))
rows = []
for th, ev in by_thread.items():
    for lab, a, b in ev:
        if not lab.startswith("scenario:"):
            continue
        mine = [(l2, x, y) for l2, x, y in ev if a <= x and y <= b and l2 != lab]
        dn = [x for l2, x, y in mine if l2 == "rng_done"]
        t_done = dn[0] if dn else b
        held = [(l2, x, y) for l2, x, y in mine if y <= t_done + 1e-9 and l2 != "rng_done"]
        rng = sum(y - x for l2, x, y in held if l2.startswith("rng:"))
        # top-level non-generator spans while the generator was held
        other = {}
        for l2, x, y in held:
            if l2.startswith("rng:"):
                continue
            nested = any(l3 != l2 and x3 <= x and y <= y3 and not l3.startswith("rng:")
                         for l3, x3, y3 in held if (l3, x3, y3) != (l2, x, y))
            if not nested:
                other[l2] = other.get(l2, 0.0) + (y - x)
        after = {}
        for l2, x, y in mine:
            if x >= t_done and l2 != "rng_done":
                nested = any(l3 != l2 and x3 <= x and y <= y3
                             for l3, x3, y3 in mine if (l3, x3, y3) != (l2, x, y) and x3 >= t_done)
                if not nested:
                    after[l2] = after.get(l2, 0.0) + (y - x)
        rows.append({"scenario": lab[9:], "got_generator": round(a, 4),
                     "released": round(t_done, 4), "end": round(b, 4),
                     "held_ms": round((t_done - a) * 1e3, 1), "in_rng_ms": round(rng * 1e3, 1),
                     "held_other_ms": {k: round(v * 1e3, 1) for k, v in other.items()},
                     "after_ms": {k: round(v * 1e3, 1) for k, v in after.items()}})
rows.sort(key=lambda r: r["got_generator"])
chunks = [(a - t_begin, b - t_begin) for th, lab, a, b in spans if lab == "chunk"]
blocks = [(a - t_begin, b - t_begin) for th, lab, a, b in spans if lab == "block"]
block_stats = {"blocks": len(blocks), "chunks": len(chunks),
               "chunk_thread_s": round(sum(b - a for a, b in chunks), 4),
               "block_wall_s": round(sum(b - a for a, b in blocks), 4),
               "chunk_ms_median": round(1e3 * float(np.median([b - a for a, b in chunks])), 2)
               if chunks else None,
               "chunk_ms_max": round(1e3 * max(b - a for a, b in chunks), 2) if chunks else None}
out = {"blocks": block_stats, "wall_s": round(wall, 4), "draws": args.draws, "config": args.config,
       "host_threads": _hostpar.N_THREADS,
       "chain_end_s": max(r["released"] for r in rows) if rows else None,
       "rng_total_s": round(sum(r["in_rng_ms"] for r in rows) / 1e3, 4),
       "held_total_s": round(sum(r["held_ms"] for r in rows) / 1e3, 4),
       "scenarios": rows}
print(json.dumps(out))
if args.json:
    with open(args.json, "w") as f:
        json.dump({"summary": out,
                   "spans": [(th, lab, round(a - t_begin, 5), round(b - t_begin, 5))
                             for th, lab, a, b in spans]}, f)
