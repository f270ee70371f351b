#!/bin/bash
mkdir -p gpurun_out
for lib in tw20 tw24 tw32 tw24d tw32d tw24g32 tw24g32d; do
  DPGICP_LIBRARY=$PWD/dpg_slam_b200/libdpgicp_$lib.so timeout 600 python tools/gpu_probe2.py corridor 5000 3,0,0 1,0,0 3,8,0 2>&1 | grep -v Warning
done > gpurun_out/probe2.log 2>&1
cat gpurun_out/probe2.log
