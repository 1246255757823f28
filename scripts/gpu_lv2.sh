#!/bin/bash
mkdir -p gpurun_out
for cc in all none; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none --cache-control $cc -k regex:locoval_tc -s 3 -c 4 --csv python scripts/lv_bench.py 2>/dev/null | grep -E "gpu__time_duration|dram__bytes_read|hit_rate" | cut -d, -f5,13- | tr '\n' ' '; echo " <- cache-control $cc"
done
