#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_final_bench_config2.json 2> gpurun_out/r2_final_bench_config2.err; echo "bench2 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_final_bench_config2.json"))
e=d["e2e"]
print("value", d["value"], d["ms_per_step"], "e2e", e["value"], round(e["ms_per_step"],1), [round(x,3) for x in e["all_walls_s"]], "engine", round(e["engine"]["ms_per_step"],1), "dev", round(e["device_sampler"]["ms_per_step"],1), "parity", d["parity"]["ok"], "cpu", d["cpu_baseline"]["value"], d["host"])
PY
