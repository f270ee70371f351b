#!/bin/bash
# Compiles the parts of the reference that are self-contained (covariance header, math_utils) from
# the sources WHERE THEY LIE under $REF_ROOT, against oracle/ref_stubs/, into oracle/_ref/.
# Test infrastructure; nothing here ships.  The ICP loop itself lives in PCL and cannot be built
# in this image (no PCL/FLANN/Eigen) — see DESIGN.md.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF_ROOT="${REF_ROOT:-/root/reference}"
[ -f "$REF_ROOT/src/icp_cov/cov_func_point_to_point.h" ] || { echo "reference not present at $REF_ROOT"; exit 1; }
mkdir -p "$HERE/_ref"
# same flags as the reference's Release build (CMakeLists.txt:9-14: -std=c++14 -O2), no FMA available on plain x86-64
/usr/bin/g++ -std=c++14 -O2 -fPIC -shared -w \
    -I"$HERE/ref_stubs" -I"$REF_ROOT/src" \
    -o "$HERE/_ref/libdpgref.so" "$HERE/ref_harness.cc" "$REF_ROOT/src/dpg_slam/math_utils.cc"
echo "built $HERE/_ref/libdpgref.so"
