#!/bin/bash
# Round-2 (second half) GPU visit: all GPU tests + smoke, bench (both arms), ncu launch list and full captures exported to CSV on the
# box (the .ncu-rep files are too large to travel back).  Usage: scripts/gpu_round3.sh [tag]
TAG=${1:-r02u}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 8 --warmup 32 --no-cpu-baseline --no-variants --no-train --locoval-batch 65536 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'physics_soa_kernel|post_step_kernel|linear_chain_kernel|linear_bf16x3_kernel|locoval_tc_kernel' \
    -s 150 -c 16 -o /tmp/${TAG}_prof python bench.py --steps 4 --warmup 8 --no-cpu-baseline --no-variants --no-train --locoval-batch 1048576 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep ${TAG}
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['segments_ms'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'])"
