#!/bin/bash
# first GPU contact: smoke (parity vs oracle) then the timing sweep
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python tools/gpu_probe.py corridor 5000 > gpurun_out/probe_corridor.log 2>&1; echo "rc=$?"
tail -50 gpurun_out/probe_corridor.log
