#!/bin/bash
# One GPU session: parity tests, bench, ncu launch list, ncu full capture of the ICP kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:icp_pairs -s 3 -c 1 -o gpurun_out/prof_icp -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
