#!/bin/bash
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-parity --no-cpu-baseline > gpurun_out/r2_bench_${label}.json 2> gpurun_out/r2_bench_${label}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${label}.json"))
e=d["e2e"]
print("${label}", "e2e_ms", round(e["ms_per_step"],1), [round(x,3) for x in e["all_walls_s"]], "engine", round(e["engine"]["ms_per_step"],1))
PY
}
run base X=1
run nice10 TRI_B200_NICE_BLOCKS=10
run nice19 TRI_B200_NICE_BLOCKS=19
run arena MALLOC_ARENA_MAX=1 TRI_B200_MALLOC_TUNE=1
run arena_nice MALLOC_ARENA_MAX=1 TRI_B200_MALLOC_TUNE=1 TRI_B200_NICE_BLOCKS=10
run base2 X=1
