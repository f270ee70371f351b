#!/bin/bash
# round 2, N GPUs of one box: multi-GPU parity tests, then the bench workloads at N (reduced sizes unless FULL=1)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt; nproc >> gpurun_out/gpus_n$N.txt; free -g >> gpurun_out/gpus_n$N.txt
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_multi_gpu.py "tests/test_gpu_round2.py::test_two_contexts_one_process_gather_equals_single_context" "tests/test_gpu_round2.py::test_cpp_runner_multi_context_writes_the_same_csv" -q -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi_n$N.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for W in ${WORKLOADS:-corridor dense multisession}; do
  EXTRA=""
  if [ "$W" = "corridor" ]; then EXTRA="--steps 10 --warmup 3"; fi
  if [ "$W" = "loop_closure" ]; then EXTRA="--steps 5 --warmup 3"; fi
  if [ "$W" = "dense" ]; then EXTRA="--steps ${DENSE_STEPS:-2} --warmup 3 ${DENSE_ARGS:---pairs 40000}"; fi
  if [ "$W" = "multisession" ]; then EXTRA="--steps 1 --warmup 3 ${MS_ARGS:---scans-per-session 4000 --world-m 28.3}"; fi
  timeout ${TMO:-1200} $TR bench.py --gpus $N --workload $W $EXTRA > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err; echo "bench $W N=$N rc=$?"
  cut -c1-300 gpurun_out/bench_${W}_n$N.json; tail -4 gpurun_out/bench_${W}_n$N.err | cut -c1-300
  if [ -n "$REFARM" ]; then
    timeout 600 $TR bench.py --impl reference --gpus $N --workload $W --steps 2 --warmup 1 > gpurun_out/bench_ref_${W}_n$N.json 2> gpurun_out/bench_ref_${W}_n$N.err; echo "ref $W rc=$?"
  fi
done
