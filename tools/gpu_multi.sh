#!/bin/bash
# multi-GPU session (N GPUs of one box): NCCL parity test + torchrun bench at N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
