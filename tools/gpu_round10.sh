#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for lib in libdpgicp_base.so libdpgicp.so; do
  DPGICP_LIBRARY=$PWD/dpg_slam_b200/$lib timeout 600 python tools/gpu_probe2.py corridor 5000 4,8,16 4 2>&1 | grep -v Warn | awk '{print $1, $2, $4, $8, $10}'
  DPGICP_LIBRARY=$PWD/dpg_slam_b200/$lib timeout 600 python tools/gpu_probe2.py loop 20000 4,8,16 2>&1 | grep -v Warn | awk '{print $1, $2, $4, $8, $10}'
done
