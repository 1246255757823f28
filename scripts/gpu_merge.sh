#!/bin/bash
TAG=${1:-m1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_rollout.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python bench.py --no-cpu-baseline --no-variants --no-train --locoval-batch 65536 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['segments_ms'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'])"
timeout 300 python bench.py --no-merge --steps 32 --no-cpu-baseline --no-variants --no-train --locoval-batch 65536 > gpurun_out/${TAG}_bench_nomerge.json 2>> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_nomerge.json')); print(d['value'], d['ms_per_step'], d['segments_ms'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'])"
tail -5 gpurun_out/${TAG}_bench.err
