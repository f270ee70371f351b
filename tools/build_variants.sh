#!/bin/bash
# development aid: build libdpgicp variants with different compile-time knobs for A/B timing on the GPU
cd "$(dirname "$0")/../dpg_slam_b200/csrc"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC,-ffp-contract=off -ccbin /usr/bin/g++ -I../../include -shared -lcudart"
build() { name=$1; shift; $NV "$@" -Xptxas -v -o ../libdpgicp_$name.so dpgicp_abi.cu 2> /tmp/ptxas_$name.txt & }
build tw20 -DDPGICP_TARGET_WARPS=20
build tw24 -DDPGICP_TARGET_WARPS=24
build tw32 -DDPGICP_TARGET_WARPS=32


build tw24g32 -DDPGICP_TARGET_WARPS=24 -DDPGICP_GROUP=32

wait
for n in tw20 tw24 tw32 tw24g32; do echo == $n; python3 - $n <<'PY'
import re,sys
t=open(f'/tmp/ptxas_{sys.argv[1]}.txt').read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info    : Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info    : Used (\d+) registers", t):
    if 'icp_pairs' in m.group(1) and 'Lb1' in m.group(1): print(' ', m.group(1)[24:33], 'spill',m.group(3),'regs',m.group(5))
PY
done
