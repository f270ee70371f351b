#!/bin/bash
# compute-sanitizer over the all-paths case (the pool refused the tool in round 1: record what happens now)
mkdir -p gpurun_out
timeout 300 python tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/sanitize_plain.log
for TOOL in memcheck racecheck synccheck; do
  timeout ${SAN_TMO:-900} compute-sanitizer --tool $TOOL --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/sanitize_$TOOL.log 2>&1; echo "$TOOL rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize case done|========= (Error|Race|Hazard|Invalid|Barrier)" gpurun_out/sanitize_$TOOL.log | head -8 | cut -c1-240
  tail -2 gpurun_out/sanitize_$TOOL.log | cut -c1-240
done
# the GPU suite once against the library built with -DDPGICP_CHECK (device-side assertions)
if [ -f dpg_slam_b200/libdpgicp_check.so ]; then
  DPGICP_LIBRARY=$PWD/dpg_slam_b200/libdpgicp_check.so timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_checked.log 2>&1; echo "checked suite rc=$?"
  tail -3 gpurun_out/pytest_checked.log | cut -c1-300; grep -c "DPGICP_CHECK failed" gpurun_out/pytest_checked.log
fi
