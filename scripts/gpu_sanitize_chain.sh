#!/bin/bash
# compute-sanitizer passes over the layer-chain tests, the merged-schedule test and the smoke run
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests/test_gpu_chain.py tests/test_gpu_rollout.py -q -k "chain or merged_schedule" > gpurun_out/sanc_${tool}_tests.log 2>&1
  echo "$tool tests exit $? :: $(grep -E 'passed|failed' gpurun_out/sanc_${tool}_tests.log | tail -1) :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanc_${tool}_tests.log | tail -1)"
  grep -E "Race reported|Invalid|Barrier error|hazard" gpurun_out/sanc_${tool}_tests.log | sort | uniq -c | head -8
done
timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanc_memcheck_smoke.log 2>&1
echo "memcheck smoke exit $? :: $(grep -E 'ERROR SUMMARY' gpurun_out/sanc_memcheck_smoke.log | tail -1)"
