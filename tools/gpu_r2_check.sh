#!/bin/bash
# round 2: correctness first (GPU parity suite + smoke), then the default bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1; nproc >> gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-400
timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err | cut -c1-400
