#!/bin/bash
# Builds differently tuned copies of the CUDA library for A/B runs on the GPU box:
#   scripts/build_variants.sh name "-DFLAG=1 ..." [source-root]
# -> scripts/lib_<name>.so (git-ignored; travels with gpurun).  Select with TRI_B200_LIB.
set -e
name=$1; flags=$2; root=${3:-$(dirname "$0")/..}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC \
     -ccbin /usr/bin/g++ $flags -o "$(dirname "$0")/lib_${name}.so" "$root/triceratops_b200/csrc/tri_cabi.cu"
echo "built scripts/lib_${name}.so ($flags)"
