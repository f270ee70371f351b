#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:icp_pairs -s 12 -c 4 -o gpurun_out/prof_icp_v6 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
