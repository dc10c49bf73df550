#!/bin/bash
# GPU box: bench.py --kernel-only for each scripts/lib_<name>.so given -> gpurun_out/ab_<tag>.jsonl
tag=$1; shift
mkdir -p gpurun_out
for name in "$@"; do
  TRI_B200_LIB=$PWD/scripts/lib_${name}.so python bench.py --kernel-only --steps 3 --warmup 2 \
      2>gpurun_out/ab_${tag}_${name}.err | tail -1 >> gpurun_out/ab_${tag}.jsonl
done
cat gpurun_out/ab_${tag}.jsonl
