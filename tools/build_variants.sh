#!/bin/bash
# development aid: build libdpgicp variants with different compile-time knobs for A/B timing on the GPU
# (tools/gpu_ab.sh <variant> ...).  usage: tools/build_variants.sh name=-DFLAG[,-DFLAG2] ...
cd "$(dirname "$0")/../dpg_slam_b200/csrc"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fopenmp -ccbin /usr/bin/g++ -I../../include -shared -lcudart -lgomp"
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}; flags=${flags//,/ }
  $NV $flags -Xptxas -v -o ../libdpgicp_$name.so dpgicp_abi.cu 2> /tmp/ptxas_$name.txt &
done
wait
for spec in "$@"; do
  name=${spec%%=*}
  echo "== $name"; grep -A2 "icp_pairs_kernelILi4ELi1ELi1EEEvNS_12KernelParamsE' for" /tmp/ptxas_$name.txt | tail -2
done
