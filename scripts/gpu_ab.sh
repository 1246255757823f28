#!/bin/bash
# A/B of an environment switch on the bench headline: scripts/gpu_ab.sh VAR
mkdir -p gpurun_out
for v in 0 1 0 1; do
  env $1=$v timeout 600 python bench.py --no-cpu-baseline --no-variants --locoval-batch 65536 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1=$v', round(d['value']), {k:round(x*1000,1) for k,x in d['segments_ms'].items()})"
done
