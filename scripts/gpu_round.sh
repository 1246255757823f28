#!/bin/bash
# One GPU visit: parity tests, smoke, bench, ncu launch list + full capture.  Usage: scripts/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
tail -3 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
# ncu launch list of the bench command (cold-cache, serialised: shares only) and one full capture of our top kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 8 --warmup 32 --no-cpu-baseline --no-variants --locoval-batch 65536 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'physics_kernel|post_step_kernel|linear_fma_kernel|locoval_kernel|linear_tc' \
    -s 200 -c 24 -o gpurun_out/${TAG}_prof python bench.py --steps 4 --warmup 8 --no-cpu-baseline --no-variants --locoval-batch 1048576 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/
