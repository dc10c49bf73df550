#!/bin/bash
# compute-sanitizer on the round-2 kernels: smoke path (geometry, lnl_kernel with its shared-memory
# draw record, finalize_kernel), the fused device sampler, the submit/wait path
mkdir -p gpurun_out
out=gpurun_out/r2_compute_sanitizer.txt
echo "compute-sanitizer (B200, round 2, final kernels)" > $out
san() { # label tool command...
  label=$1; tool=$2; shift 2
  timeout 500 compute-sanitizer --tool $tool "$@" > gpurun_out/_san.log 2>&1
  echo "$label | $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/_san.log | tail -1) | $(grep -E '[0-9]+ passed|smoke ok|failed' gpurun_out/_san.log | tail -1)" | tee -a $out
}
for tool in memcheck racecheck initcheck; do
  san "smoke()" $tool python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
done
for tool in memcheck racecheck; do
  san "tests/test_gpu_async.py" $tool python -m pytest tests/test_gpu_async.py -x -q
done
san "device sampler (test_device_sampler.py -k global_draw_index,eccentricity)" memcheck python -m pytest tests/test_device_sampler.py -x -q -k "global_draw_index or eccentricity"
san "scalar loop + per-point sigma" memcheck python -m pytest tests/test_gpu_pointwise_sigma.py -x -q
cat $out
