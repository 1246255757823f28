#!/bin/bash
# Round-2 GPU visit: tests, smoke, bench (both arms), ncu launch lists + full captures.  Usage: scripts/gpu_round2.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
head -c 400 gpurun_out/${TAG}_bench_ref.json; echo
# ncu launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 8 --warmup 32 --no-cpu-baseline --no-variants --no-train --locoval-batch 65536 > gpurun_out/${TAG}_ncu_bench.log 2>&1
# full captures: rollout kernels, then the update step's kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'physics_soa_kernel|post_step_kernel|linear_bf16x3_kernel|locoval_tc_kernel' \
    -s 200 -c 24 -o gpurun_out/${TAG}_prof python bench.py --steps 4 --warmup 8 --no-cpu-baseline --no-variants --no-train --locoval-batch 1048576 > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'xform2_kernel|linear_bf16x3_kernel|adam_clip_kernel|column_moments_kernel' \
    -s 120 -c 30 -o gpurun_out/${TAG}_prof_update python scripts/profile_update.py 16384 1 > gpurun_out/${TAG}_ncu_update.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
