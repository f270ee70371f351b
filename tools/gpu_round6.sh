#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --no-cpu-baseline --workload loop_closure > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err; echo rc=$?; cut -c1-300 gpurun_out/bench_loop.json; tail -3 gpurun_out/bench_loop.err
