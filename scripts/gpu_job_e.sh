#!/bin/bash
# 8-GPU: dist_check, weak + strong bench (config 2), weak config 4
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 scripts/dist_check.py --repeats 2 > gpurun_out/r2_dist_check_${N}gpu.json 2> gpurun_out/r2_dist_check_${N}gpu.err
echo "dist_check rc=$?"; tail -c 1200 gpurun_out/r2_dist_check_${N}gpu.json; tail -3 gpurun_out/r2_dist_check_${N}gpu.err
timeout 400 $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_weak.json 2> gpurun_out/r2_bench_${N}gpu_weak.err
echo "weak rc=$?"; cut -c1-2200 gpurun_out/r2_bench_${N}gpu_weak.json; tail -3 gpurun_out/r2_bench_${N}gpu_weak.err
timeout 400 $TR --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong > gpurun_out/r2_bench_${N}gpu_strong.json 2> gpurun_out/r2_bench_${N}gpu_strong.err
echo "strong rc=$?"; cut -c1-2200 gpurun_out/r2_bench_${N}gpu_strong.json; tail -3 gpurun_out/r2_bench_${N}gpu_strong.err
timeout 400 $TR --master-port 29524 bench.py --gpus $N --steps 5 --warmup 3 --config 4 > gpurun_out/r2_bench_${N}gpu_config4.json 2> gpurun_out/r2_bench_${N}gpu_config4.err
echo "config4 rc=$?"; cut -c1-1500 gpurun_out/r2_bench_${N}gpu_config4.json; tail -3 gpurun_out/r2_bench_${N}gpu_config4.err
