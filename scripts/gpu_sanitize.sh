#!/bin/bash
# compute-sanitizer passes over the smoke run (physics, post-step, one rollout step incl. the tcgen05 layers) and the LocoVal /
# trajectory-reset / fine-tuning tests; logs to gpurun_out/san_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_${tool}_smoke.log 2>&1
  echo "$tool smoke exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/san_${tool}_smoke.log | head -8
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_traj_reset.py tests/test_gpu_parity.py -m gpu -q -x \
      -k "traj_reset_matches or finetune_step or multimodal or philox" > gpurun_out/san_${tool}_tests.log 2>&1
  echo "$tool tests exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|passed|failed" gpurun_out/san_${tool}_tests.log | head -8
done
