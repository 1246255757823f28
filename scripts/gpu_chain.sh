#!/bin/bash
# GPU visit for the layer chain: its tests first (under a short timeout: a scheduling bug would hang), then rollout tests, bench A/B
TAG=${1:-c1}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q > gpurun_out/${TAG}_chain.log 2>&1; rc=$?; echo "chain pytest exit $rc" >> gpurun_out/${TAG}_chain.log
tail -15 gpurun_out/${TAG}_chain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests/test_gpu_rollout.py -x -q > gpurun_out/${TAG}_rollout.log 2>&1; echo "rollout pytest exit $?" >> gpurun_out/${TAG}_rollout.log
tail -5 gpurun_out/${TAG}_rollout.log
timeout 300 python bench.py --no-cpu-baseline --no-variants --no-train --locoval-batch 65536 > gpurun_out/${TAG}_bench_chain.json 2> gpurun_out/${TAG}_bench_chain.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_chain.json')); print(d['value'], d['ms_per_step'], d['segments_ms'], d['e2e']['value'], d['gpu_launches'])"
timeout 300 python bench.py --no-chain --no-cpu-baseline --no-variants --no-train --locoval-batch 65536 > gpurun_out/${TAG}_bench_nochain.json 2> gpurun_out/${TAG}_bench_nochain.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_nochain.json')); print(d['value'], d['ms_per_step'], d['segments_ms'], d['e2e']['value'], d['gpu_launches'])"
tail -3 gpurun_out/${TAG}_bench_chain.err
