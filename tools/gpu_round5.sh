#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
DPGICP_LIBRARY=$PWD/dpg_slam_b200/libdpgicp_stats.so timeout 600 python tools/gpu_probe2.py corridor 5000 3,0,0 2>&1 | grep -v Warning
DPGICP_LIBRARY=$PWD/dpg_slam_b200/libdpgicp_stats.so timeout 600 python tools/gpu_probe2.py loop 20000 3,0,0 2>&1 | grep -v Warning
timeout 600 python tools/gpu_probe2.py loop 20000 3,0,0 1,0,0 2>&1 | grep -v Warning
