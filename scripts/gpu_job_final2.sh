#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_final_bench_config2.json 2> gpurun_out/r2_final_bench_config2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --config 3 --no-cpu-baseline > gpurun_out/r2_final_bench_config3.json 2> gpurun_out/r2_final_bench_config3.err; echo "bench3 rc=$?"
python - <<PY
import json
for c in (2,3):
    d=json.load(open("gpurun_out/r2_final_bench_config%d.json"%c))
    e=d["e2e"]
    print(c, "value", d["value"], d["ms_per_step"], "e2e", e["value"], round(e["ms_per_step"],1), [round(x,3) for x in e["all_walls_s"]], "engine", round(e["engine"]["ms_per_step"],1), "dev", round(e["device_sampler"]["ms_per_step"],1), "parity", d["parity"]["ok"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
timeout 200 python -m pytest tests/test_gpu_lnz.py tests/test_bufpool.py -x -q 2>&1 | tail -1
timeout 200 python scripts/chain_trace.py > gpurun_out/r2_chain_h.json 2>/dev/null; cut -c1-330 gpurun_out/r2_chain_h.json
