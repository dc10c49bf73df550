#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_host_blocks.py tests/test_host_blocks_e2e.py -x -q 2>&1 | tail -2
python scripts/host_only.py > gpurun_out/r2_host_only.json 2>/dev/null; cut -c1-900 gpurun_out/r2_host_only.json
TRI_B200_SCENARIO_THREADS=1 python scripts/host_only.py > gpurun_out/r2_host_only_st1.json 2>/dev/null; cut -c1-200 gpurun_out/r2_host_only_st1.json
timeout 600 python bench.py > gpurun_out/r2_final_bench_config2.json 2> gpurun_out/r2_final_bench_config2.err; echo "bench2 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_final_bench_config2.json"))
e=d["e2e"]
print("value", d["value"], d["ms_per_step"], "e2e", e["value"], round(e["ms_per_step"],1), [round(x,3) for x in e["all_walls_s"]], "engine", round(e["engine"]["ms_per_step"],1), "dev", round(e["device_sampler"]["ms_per_step"],1), "parity", d["parity"]["ok"], "cpu", d["cpu_baseline"]["value"])
PY
timeout 200 python scripts/chain_trace.py > gpurun_out/r2_chain_h.json 2>/dev/null; cut -c1-330 gpurun_out/r2_chain_h.json
