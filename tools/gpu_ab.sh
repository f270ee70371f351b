#!/bin/bash
# development aid: parity tests, then A/B kernel timing of library variants on one box
# usage: tools/gpu_ab.sh variant1 variant2 ...   (libdpgicp_<variant>.so; "default" = libdpgicp.so)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python - <<'PY'
from dpg_slam_b200.scanmatch import ScanMatcher
with ScanMatcher(0) as sm: print("fp32 probe", sm.fp32_probe())
PY
for v in "$@"; do
  if [ "$v" = default ]; then unset DPGICP_LIBRARY; else export DPGICP_LIBRARY=$PWD/dpg_slam_b200/libdpgicp_$v.so; fi
  timeout 600 python tools/gpu_probe2.py corridor 5000 default 2>&1 | grep -v Warning | tail -3
  timeout 600 python tools/gpu_probe2.py loop 20000 default 2>&1 | grep -v Warning | tail -3
done
