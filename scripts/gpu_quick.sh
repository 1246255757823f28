#!/bin/bash
# quick GPU visit: all gpu tests (no -x) + bench (tensor cores)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --tensor-cores > gpurun_out/${TAG}_bench_tc.json 2> gpurun_out/${TAG}_bench_tc.err; echo "bench tc exit $?"; cat gpurun_out/${TAG}_bench_tc.json; tail -3 gpurun_out/${TAG}_bench_tc.err
