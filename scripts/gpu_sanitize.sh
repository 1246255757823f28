#!/bin/bash
# compute-sanitizer passes (memcheck, racecheck, synccheck) over the smoke run and a spread of the GPU tests; logs to
# gpurun_out/san_*.log, one summary line per pass on stdout
mkdir -p gpurun_out
SEL=${1:-"traj_reset_matches or finetune_step or multimodal or philox or lockstep or first_steps or fused_head or split_output or single_env_step or fused_sinks or value_reuse or host_observation or gym_shim or post_step_matches or locoval_tensor_core"}
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_${tool}_smoke.log 2>&1
  echo "$tool smoke exit $? :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}_smoke.log | tail -1)"
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests -m gpu -q -k "$SEL" > gpurun_out/san_${tool}_tests.log 2>&1
  echo "$tool tests exit $? :: $(grep -E 'passed|failed' gpurun_out/san_${tool}_tests.log | tail -1) :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}_tests.log | tail -1)"
  grep -E "Race reported|Invalid|Barrier error|hazard" gpurun_out/san_${tool}_tests.log | sort | uniq -c | head -8
done
