#!/bin/bash
# Round-2 GPU visit: bench (both arms), ncu launch lists + full captures exported to CSV on the box (the .ncu-rep files are too
# large to travel back).  Usage: scripts/gpu_round2.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 8 --warmup 32 --no-cpu-baseline --no-variants --no-train --locoval-batch 65536 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'physics_soa_kernel|post_step_kernel|linear_bf16x3_kernel|locoval_tc_kernel' \
    -s 200 -c 24 -o /tmp/${TAG}_prof python bench.py --steps 4 --warmup 8 --no-cpu-baseline --no-variants --no-train --locoval-batch 1048576 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:'xform2_kernel|linear_bf16x3_kernel|adam_clip_kernel|column_moments_kernel' \
    -s 120 -c 30 -o /tmp/${TAG}_prof_update python scripts/profile_update.py 16384 1 > gpurun_out/${TAG}_ncu_update.log 2>&1
ncu -i /tmp/${TAG}_prof_update.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_update_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep ${TAG}
