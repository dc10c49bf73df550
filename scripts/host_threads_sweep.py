"""calc_probs wall time (host sampler, N = 1e6) against TRI_B200_SCENARIO_THREADS on this box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, time, json, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch, _workloads
from triceratops_b200.engine import get_engine
get_engine(0)
lc = _workloads.lightcurve(2)
tgt = _workloads.make_target(2)
walls = []
for _ in range(5):
    t0 = time.perf_counter(); _workloads.run_calc_probs(tgt, 2, lc, 1_000_000, 2026); torch.cuda.synchronize()
    walls.append(time.perf_counter() - t0)
print(json.dumps({"threads": os.environ.get("TRI_B200_SCENARIO_THREADS"), "walls": walls, "FPP": float(tgt.FPP)}))
''' % (ROOT, ROOT)
for th in sys.argv[1:] or ["4", "6", "8", "12"]:
    env = dict(os.environ, TRI_B200_SCENARIO_THREADS=th)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:])
