#!/bin/bash
mkdir -p gpurun_out
TRI_B200_RNG_TRACE=1 python scripts/host_rng_bench.py > gpurun_out/r2_host_rng.json 2>gpurun_out/r2_host_rng.err; cat gpurun_out/r2_host_rng.json; grep "beta n=1000000" gpurun_out/r2_host_rng.err | tail -4
python scripts/chain_trace.py 2>/dev/null | tail -1 | tee gpurun_out/r2_chain_trace.json
python -m pytest tests/test_device_sampler.py tests/test_gpu_pointwise_sigma.py -m gpu -q 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -2 gpurun_out/r2_bench_c.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c.json')); e=d['e2e']
print('value',d['value'],d['ms_per_step'],'e2e',e['value'],e['ms_per_step'],'engine',e['engine']['ms_per_step'],'device',e['device_sampler']['ms_per_step'],'parity',d['parity']['ok'])"
