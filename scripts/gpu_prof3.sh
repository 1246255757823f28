#!/bin/bash
# evidence run: launch list (eager so that every launch is visible to ncu) + full capture of the main kernels
TAG=${1:-r01c}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --eager --serial --steps 8 --warmup 8 --no-cpu-baseline --no-variants --locoval-batch 65536 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'linear_bf16x3_kernel|physics_soa_kernel|post_step_kernel|locoval' -s 140 -c 40 \
    -o gpurun_out/${TAG}_prof python bench.py --eager --serial --steps 4 --warmup 4 --no-cpu-baseline --no-variants --locoval-batch 1048576 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json | cut -c1-300
