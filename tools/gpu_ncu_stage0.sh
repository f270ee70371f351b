#!/bin/bash
# development aid: source-level counters of the first-stage launch of one bench step -> gpurun_out/prof_stage0_<tag>.ncu-rep
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section InstructionStats --section SpeedOfLight --section LaunchStats --section Occupancy \
  --clock-control none --import-source on -k regex:icp_pairs -s 12 -c 1 -o gpurun_out/prof_stage0_$TAG -f $CMD > gpurun_out/ncu_stage0.log 2>&1
echo "ncu stage0 rc=$?"
