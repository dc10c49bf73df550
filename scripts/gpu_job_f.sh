#!/bin/bash
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" timeout 200 python scripts/chain_trace.py --json gpurun_out/r2_chain_${label}_full.json > gpurun_out/r2_chain_${label}.json 2> gpurun_out/r2_chain_${label}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_chain_${label}.json"))
print("${label}", "wall", d["wall_s"], "chain_end", d["chain_end_s"], "rng", d["rng_total_s"], "held", d["held_total_s"], "threads", d["host_threads"], d["blocks"])
PY
}
run v3a X=1
run v3b X=1
run v3st6 TRI_B200_SCENARIO_THREADS=6
TRI_B200_RNG_TRACE=1 python - <<'PY' 2>&1 | tail -4
import numpy as np, time
from triceratops_b200 import _fastrng
np.random.seed(1)
for _ in range(4):
    t0=time.perf_counter(); _fastrng.beta_rvs(0.867,3.03,1_000_000); print("beta ms", round((time.perf_counter()-t0)*1e3,2))
for f,name in ((lambda:_fastrng.rand(1_000_000),"rand"),(lambda:_fastrng.randint(0,4321,1_000_000),"randint"),(lambda:_fastrng.skip(1_000_000),"skip")):
    b=1e9
    for _ in range(9):
        t0=time.perf_counter(); f(); b=min(b,time.perf_counter()-t0)
    print(name, round(b*1e3,2))
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-parity --cpu-draws 2000 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_f.json')); print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','all_walls_s')}, d['e2e']['engine']['ms_per_step'], d['e2e']['device_sampler']['ms_per_step'])"; tail -3 gpurun_out/r2_bench_f.err
