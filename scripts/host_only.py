"""Steady-state wall time of calc_probs' HOST side alone (prior draws in numpy's stream,
element-wise preparation, column assembly) with a recording stand-in for the engine: the part
of the public call that no GPU can shorten.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _workloads  # noqa: E402
from triceratops_b200 import _blocks, _hostpar  # noqa: E402

config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
lc = _workloads.lightcurve(config)
walls = []
for _ in range(6):
    walls.append(_workloads.record_calls(config, N, 2026, lc)[1])
spans = []
real = _blocks.run


def timed(kind, n, **kw):
    t0 = time.perf_counter()
    out = real(kind, n, **kw)
    spans.append((kind, out is not None, time.perf_counter() - t0))
    return out


_blocks.run = timed
import triceratops_b200.marginal_likelihoods as ml  # noqa: E402
ml._blocks.run = timed
w = _workloads.record_calls(config, N, 2026, lc)[1]
print(json.dumps({"config": config, "draws": N, "host_threads": _hostpar.N_THREADS,
                  "host_only_walls_s": [round(x, 4) for x in walls], "traced_wall_s": round(w, 4),
                  "c_blocks": [(k, ok, round(t * 1e3, 2)) for k, ok, t in spans],
                  "c_block_total_ms": round(sum(t for _, _, t in spans) * 1e3, 1)}))
