#!/bin/bash
# ncu --set full with source correlation for ONE kernel (regex $1), launch skip $2; writes the SASS/source page as CSV
K=${1:-post_step_kernel}; SKIP=${2:-40}; TAG=${3:-src}
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$K" -s $SKIP -c 1 -o /tmp/${TAG} python bench.py --steps 4 --warmup 8 --no-cpu-baseline --no-variants --no-train --locoval-batch 65536 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sass.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*; head -c 600 gpurun_out/${TAG}_sass.csv
