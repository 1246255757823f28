#!/bin/bash
# chain tests (short timeout) + trace + bench
TAG=${1:-c2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q > gpurun_out/${TAG}_chain.log 2>&1; rc=$?; echo "chain pytest exit $rc" >> gpurun_out/${TAG}_chain.log
tail -15 gpurun_out/${TAG}_chain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python scripts/chain_trace.py ${TAG} > gpurun_out/${TAG}_trace.log 2>&1; echo "trace exit $?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_chain_trace.json')); print({k:v for k,v in d.items() if k!='single'})"
timeout 300 python bench.py --no-cpu-baseline --no-variants --no-train --locoval-batch 65536 > gpurun_out/${TAG}_bench_chain.json 2> gpurun_out/${TAG}_bench_chain.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_chain.json')); print(d['value'], d['ms_per_step'], d['segments_ms'], d['e2e']['value'], d['gpu_launches'])"
tail -3 gpurun_out/${TAG}_bench_chain.err
