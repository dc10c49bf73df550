#!/bin/bash
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" timeout 200 python scripts/chain_trace.py > gpurun_out/r2_chain_${label}.json 2> gpurun_out/r2_chain_${label}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_chain_${label}.json"))
print("${label}", "wall", d["wall_s"], "chain_end", d["chain_end_s"], "rng", d["rng_total_s"], "held", d["held_total_s"], "threads", d["host_threads"], d["blocks"])
PY
}
run ch4 TRI_B200_BLOCK_CHUNKS=4
run ch6 TRI_B200_BLOCK_CHUNKS=6
run ch4b TRI_B200_BLOCK_CHUNKS=4 TRI_B200_SWITCH_INTERVAL=0.00005
run ch6b TRI_B200_BLOCK_CHUNKS=6 TRI_B200_SWITCH_INTERVAL=0.00005
run ch3 TRI_B200_BLOCK_CHUNKS=3
run ch6st6 TRI_B200_BLOCK_CHUNKS=6 TRI_B200_SCENARIO_THREADS=6
run ch4st8 TRI_B200_BLOCK_CHUNKS=4 TRI_B200_SCENARIO_THREADS=8
run ch16b TRI_B200_SWITCH_INTERVAL=0.00005
