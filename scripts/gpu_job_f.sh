#!/bin/bash
mkdir -p gpurun_out
python scripts/host_only.py > gpurun_out/r2_host_only.json 2> gpurun_out/r2_host_only.err; cat gpurun_out/r2_host_only.json; tail -3 gpurun_out/r2_host_only.err
TRI_B200_SCENARIO_THREADS=1 python scripts/host_only.py > gpurun_out/r2_host_only_st1.json 2>/dev/null; cat gpurun_out/r2_host_only_st1.json
MALLOC_ARENA_MAX=1 TRI_B200_MALLOC_TUNE=1 python scripts/host_only.py > gpurun_out/r2_host_only_arena.json 2>/dev/null; cat gpurun_out/r2_host_only_arena.json
