#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16x3" > gpurun_out/cta2_pytest.log 2>&1; echo "exit $?" >> gpurun_out/cta2_pytest.log
tail -25 gpurun_out/cta2_pytest.log
nvidia-smi --query-gpu=name,memory.used --format=csv
