#!/bin/bash
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" timeout 200 python scripts/chain_trace.py > gpurun_out/r2_chain_${label}.json 2> gpurun_out/r2_chain_${label}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_chain_${label}.json"))
b=d["blocks"]
print("${label}", "wall", d["wall_s"], "chain_end", d["chain_end_s"], "rng", d["rng_total_s"], "threads", d["host_threads"], "cblocks", b.get("c_blocks"), b.get("c_block_wall_s"))
PY
}
run base X=1
run bt8 TRI_B200_BLOCK_THREADS=8
run bt4 TRI_B200_BLOCK_THREADS=4
run bt12 TRI_B200_BLOCK_THREADS=12
run ht8 TRI_B200_HOST_THREADS=8
run ht12 TRI_B200_HOST_THREADS=12
run ht8bt4 TRI_B200_HOST_THREADS=8 TRI_B200_BLOCK_THREADS=4
run base2 X=1
run passive OMP_WAIT_POLICY=passive
run passive_bt8 OMP_WAIT_POLICY=passive TRI_B200_BLOCK_THREADS=8
