#!/bin/bash
# ncu launch list of the bench (graphs off so every launch is visible) + full capture of the named kernels
TAG=${1:-p}; KREGEX=${2:-linear_bf16x3_kernel|physics_kernel|post_step_kernel}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --eager --steps 8 --warmup 8 --no-cpu-baseline --no-variants --locoval-batch 65536 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -s 150 -c 30 \
    -o gpurun_out/${TAG}_prof python bench.py --eager --steps 4 --warmup 4 --no-cpu-baseline --no-variants --locoval-batch 65536 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/ | grep ${TAG}
