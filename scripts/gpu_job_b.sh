#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_gputest6.log; cat gpurun_out/r2_gputest6.log
scripts/ab_bench.sh f noacos acos
python scripts/chain_trace.py 2>/dev/null | tail -1 | tee gpurun_out/r2_chain_trace.json
