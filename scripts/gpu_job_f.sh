#!/bin/bash
mkdir -p gpurun_out
python scripts/host_only.py > gpurun_out/r2_host_only.json 2> gpurun_out/r2_host_only.err; cat gpurun_out/r2_host_only.json; tail -3 gpurun_out/r2_host_only.err
TRI_B200_SCENARIO_THREADS=1 python scripts/host_only.py > gpurun_out/r2_host_only_st1.json 2>/dev/null; cat gpurun_out/r2_host_only_st1.json
timeout 300 python bench.py --steps 6 --warmup 3 --no-parity --no-cpu-baseline > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_g.json"))
e=d["e2e"]
print("e2e_ms", round(e["ms_per_step"],1), [round(x,3) for x in e["all_walls_s"]], "engine", round(e["engine"]["ms_per_step"],1), "dev", round(e["device_sampler"]["ms_per_step"],1))
PY
timeout 200 python scripts/chain_trace.py --json gpurun_out/r2_chain_h_full.json > gpurun_out/r2_chain_h.json 2>/dev/null; cut -c1-400 gpurun_out/r2_chain_h.json
