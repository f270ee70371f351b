#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python tools/gpu_probe2.py corridor 5000 4 4,8 4,8,16 4,8,32 4,32 4,8,16,32 4,16,32 4,6,12,32 5,8,32 6,12,32 8,32 2>&1 | grep -v Warn | awk '{print $2, $4, $8}'
