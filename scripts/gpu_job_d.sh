#!/bin/bash
# 2-GPU validation: dist_check (scatter path vs rank-0 alone), weak + strong bench, reference arm idle ranks
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/r2_dist_check_2gpu.json 2> gpurun_out/r2_dist_check_2gpu.err
echo "dist_check rc=$?"; tail -c 1500 gpurun_out/r2_dist_check_2gpu.json; tail -5 gpurun_out/r2_dist_check_2gpu.err
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu_weak.json 2> gpurun_out/r2_bench_2gpu_weak.err
echo "weak rc=$?"; cat gpurun_out/r2_bench_2gpu_weak.json; tail -5 gpurun_out/r2_bench_2gpu_weak.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong > gpurun_out/r2_bench_2gpu_strong.json 2> gpurun_out/r2_bench_2gpu_strong.err
echo "strong rc=$?"; cat gpurun_out/r2_bench_2gpu_strong.json; tail -5 gpurun_out/r2_bench_2gpu_strong.err

