#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_sweep_config5.jsonl
run() { # workers threads sampler tois
  TRI_B200_HOST_THREADS=$2 TRI_B200_SCENARIO_THREADS=${5:-1} timeout 400 python scripts/sweep_config5.py --tois $4 --draws 1000000 --workers-per-gpu $1 --sampler $3 2>/dev/null | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); d['host_threads_per_worker']=$2; d['scenario_threads']=${5:-1}
print(json.dumps(d))" | tee -a gpurun_out/r2_sweep_config5.jsonl | cut -c100-420
}
run 16 1 host 64
run 8 2 host 64
run 4 4 host 64 2
run 2 8 host 48 4
run 1 16 host 32 4
run 2 4 device 96
