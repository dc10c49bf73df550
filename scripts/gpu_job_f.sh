#!/bin/bash
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-parity --no-cpu-baseline > gpurun_out/r2_bench_${label}.json 2> gpurun_out/r2_bench_${label}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${label}.json"))
e=d["e2e"]
print("${label}", "e2e_ms", round(e["ms_per_step"],1), [round(x,3) for x in e["all_walls_s"]])
PY
}
run st1 TRI_B200_SCENARIO_THREADS=1
run st2 TRI_B200_SCENARIO_THREADS=2
run st2_bt8 TRI_B200_SCENARIO_THREADS=2 TRI_B200_BLOCK_THREADS=8
run st4 X=1
