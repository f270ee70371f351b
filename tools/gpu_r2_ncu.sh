#!/bin/bash
# round 2: per-launch list of one bench step + `ncu --set full` capture of the step's icp_pairs_kernel launches
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-latency ${BENCH_ARGS:-}"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
cut -c1-200 gpurun_out/plain_$TAG.log
timeout 900 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:icp_pairs -s 12 -c 4 -o gpurun_out/prof_icp_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof_icp_$TAG.ncu-rep
