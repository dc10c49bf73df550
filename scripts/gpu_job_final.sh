#!/bin/bash
# final single-GPU validation of the round: GPU test suite, smoke, bench lines for configs 2, 3, 4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_final_bench_config2.json 2> gpurun_out/r2_final_bench_config2.err; echo "bench2 rc=$?"; cut -c1-1500 gpurun_out/r2_final_bench_config2.json; tail -2 gpurun_out/r2_final_bench_config2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/r2_final_bench_reference.json
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_config3.json 2> gpurun_out/r2_final_bench_config3.err; echo "bench3 rc=$?"; cut -c1-300 gpurun_out/r2_final_bench_config3.json
timeout 900 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r2_final_bench_config4.json 2> gpurun_out/r2_final_bench_config4.err; echo "bench4 rc=$?"; cut -c1-300 gpurun_out/r2_final_bench_config4.json
python scripts/host_rng_bench.py > gpurun_out/r2_final_host_rng.json 2>/dev/null; cut -c1-600 gpurun_out/r2_final_host_rng.json
