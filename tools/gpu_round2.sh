#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for lib in libdpgicp.so libdpgicp_g32.so libdpgicp_g8.so; do
  DPGICP_LIBRARY=$PWD/dpg_slam_b200/$lib timeout 900 python tools/gpu_probe.py corridor 5000 > gpurun_out/probe_$lib.log 2>&1; echo "probe $lib rc=$?"
done
tail -4 gpurun_out/probe_libdpgicp.so.log
