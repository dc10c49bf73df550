#!/bin/bash
# GPU box: host RNG timing, bench configs 2/3/4, ncu captures (results under gpurun_out/)
mkdir -p gpurun_out
python scripts/host_rng_bench.py > gpurun_out/r2_host_rng.json 2>gpurun_out/r2_host_rng.err
cat gpurun_out/r2_host_rng.json; tail -3 gpurun_out/r2_host_rng.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -3 gpurun_out/r2_bench_b.err
python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r2_bench_config3.json 2> gpurun_out/r2_bench_config3.err
tail -3 gpurun_out/r2_bench_config3.err
for c in 2 3; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:lnl_kernel -s 12 -c 2 \
    -o gpurun_out/r2_prof_config$c python bench.py --config $c --kernel-only --steps 1 --warmup 1 \
    > gpurun_out/r2_ncu_config$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lnl_kernel -s 12 -c 2 \
  -o gpurun_out/r2_prof_config4 python bench.py --config 4 --draws 1000000 --kernel-only --steps 1 --warmup 1 \
  > gpurun_out/r2_ncu_config4.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 60 --csv \
  --log-file gpurun_out/r2_launches.csv python bench.py --kernel-only --steps 1 --warmup 1 > gpurun_out/r2_launches.log 2>&1
timeout 1200 python bench.py --config 4 --steps 2 --warmup 3 --cpu-draws 4000 > gpurun_out/r2_bench_config4.json 2> gpurun_out/r2_bench_config4.err
tail -3 gpurun_out/r2_bench_config4.err
ls -la gpurun_out | tail -20
