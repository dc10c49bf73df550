#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for st in 1 2; do
TRI_B200_SCENARIO_THREADS=$st timeout 200 $TR --master-port 2951$st scripts/dist_check.py --repeats 6 > gpurun_out/r2_dist_check_2gpu_st$st.json 2> gpurun_out/r2_dist_check_2gpu_st$st.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_dist_check_2gpu_st$st.json"))
print("st$st", [round(x,3) for x in d["host_sampler_walls_s"]], d["all_ranks_equal"], d["lnZ_max_abs_vs_single_rank"])
PY
done
