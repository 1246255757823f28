#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bf16x3" > gpurun_out/tc_pytest.log 2>&1; echo "tc pytest exit $?" >> gpurun_out/tc_pytest.log
tail -30 gpurun_out/tc_pytest.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/all_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/all_pytest.log
tail -40 gpurun_out/all_pytest.log
timeout 600 python bench.py > gpurun_out/bench_fma.json 2> gpurun_out/bench_fma.err; echo "bench fma exit $?"; cat gpurun_out/bench_fma.json; tail -3 gpurun_out/bench_fma.err
timeout 600 python bench.py --tensor-cores --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench tc exit $?"; cat gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
