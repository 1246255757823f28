#!/bin/bash
TAG=${1:-p2}; KREGEX=${2:-physics_soa_kernel}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_physics.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -s 12 -c 3 \
    -o gpurun_out/${TAG}_prof python bench.py --eager --steps 4 --warmup 4 --no-cpu-baseline --no-variants --locoval-batch 65536 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
