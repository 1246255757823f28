#!/bin/bash
# LocoVal iteration: parity tests + 1 M-batch timing (no L2 flush: the 424 MB of inputs exceed the L2) + optional ncu capture
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "locoval" > gpurun_out/lv_pytest.log 2>&1; echo "exit $?" >> gpurun_out/lv_pytest.log
tail -30 gpurun_out/lv_pytest.log | cut -c1-250
timeout 300 python scripts/lv_bench.py
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:locoval_tc -s 3 -c 1 -o gpurun_out/lv_prof python scripts/lv_bench.py > gpurun_out/lv_ncu.log 2>&1
  tail -2 gpurun_out/lv_ncu.log
fi
