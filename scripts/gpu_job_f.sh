#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bufpool.py tests/test_gpu_lnz.py -x -q 2>&1 | tail -2
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${label}.json 2> gpurun_out/r2_bench_${label}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${label}.json"))
e=d["e2e"]
print("${label}", "e2e_ms", round(e["ms_per_step"],1), [round(x,3) for x in e["all_walls_s"]], "engine", round(e["engine"]["ms_per_step"],1), "parity", d.get("parity",{}).get("ok"))
PY
}
run pinned X=1
run nopool TRI_B200_PINNED_POOL_MB=0
run pinned2 X=1
run nopool2 TRI_B200_PINNED_POOL_MB=0
