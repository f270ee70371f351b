#!/bin/bash
# development aid: A/B kernel timing of library variants on one box
# usage: tools/gpu_ab.sh variant1 variant2 ...   (libdpgicp_<variant>.so; "default" = libdpgicp.so)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset DPGICP_LIBRARY; else export DPGICP_LIBRARY=$PWD/dpg_slam_b200/libdpgicp_$v.so; fi
  timeout 600 python tools/gpu_probe2.py corridor 5000 default 2>&1 | grep -v Warning | tail -2
  timeout 600 python tools/gpu_probe2.py loop 20000 default 2>&1 | grep -v Warning | tail -2
done
