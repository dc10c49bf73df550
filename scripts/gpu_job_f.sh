#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/chain_trace.py --json gpurun_out/r2_chain_trace_full.json > gpurun_out/r2_chain_trace.json 2> gpurun_out/r2_chain_trace.err
echo "trace rc=$?"; cat gpurun_out/r2_chain_trace.json; tail -5 gpurun_out/r2_chain_trace.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-parity --cpu-draws 2000 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
echo "bench rc=$?"; cut -c1-1800 gpurun_out/r2_bench_f.json; tail -5 gpurun_out/r2_bench_f.err
