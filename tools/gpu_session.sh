#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
./dpg_slam_b200/dpg_batch_runner --scans 801 --passes 2 --cov-mode 2 --out gpurun_out/runner.csv > gpurun_out/runner.json 2> gpurun_out/runner.err; echo "runner rc=$?"; cat gpurun_out/runner.json; head -3 gpurun_out/runner.csv
./dpg_slam_b200/dpg_batch_runner --synthetic office --scans 400 --metric 1 --divisor 1 --cov-mode 2 --write-log gpurun_out/office.scanlog >> gpurun_out/runner.json 2>> gpurun_out/runner.err; echo "runner2 rc=$?"
./dpg_slam_b200/dpg_batch_runner --log gpurun_out/office.scanlog --metric 1 --divisor 1 --cov-mode 2 >> gpurun_out/runner.json 2>> gpurun_out/runner.err; echo "runner3 rc=$?"; tail -2 gpurun_out/runner.json; rm -f gpurun_out/office.scanlog
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:icp_pairs -s 12 -c 4 -o gpurun_out/prof_icp_final -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
