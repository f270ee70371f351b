#!/bin/bash
# development aid: stage-chain and hand-over sweep on the corridor workload (kernel time of one step)
mkdir -p gpurun_out
for H in "" "1,1" "3,1" "2,2" "1.5,1.5" "4,1"; do
  for CH in default "4,16,16x4" "4,8,16x4" "4,6,10,16x4" "4,8,16,16x2" "4,8,16"; do
    if [ -z "$H" ]; then unset DPGICP_HANDOVER; else export DPGICP_HANDOVER=$H; fi
    echo "== handover '${H}' chain $CH"
    timeout 300 python tools/gpu_probe2.py corridor 5000 "$CH" 2>&1 | grep "_d1" | sed 's/"evals.*//' 
  done
done
