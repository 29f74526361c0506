#!/usr/bin/env bash
# Builds librtx_b200.so (the C-ABI library, sm_100a only) in-tree, next to this script.
#   -lineinfo                 source mapping for ncu
#   -Xcompiler -ffp-contract=off   host doubles (wall basis, Camera::init) round like the reference's x86-64 build
# Device double arithmetic on the parity path uses explicit __d*_rn intrinsics, so -fmad stays at its default
# for the FP32 screen's FFMAs.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-ffp-contract=off,-Wall ${RTX_NVCC_EXTRA:-} \
    -I"$here/../include" -I"$here/csrc" -shared \
    "$here/csrc/api.cu" "$here/csrc/trace.cu" "$here/csrc/trace_grid.cu" "$here/csrc/aux_kernels.cu" "$here/csrc/tonemap.cu" \
    -o "$here/librtx_b200.so"
echo "built $here/librtx_b200.so"
# Test-only builds with 1-slot internal buffers (tests/test_gpu_parity.py::test_overflow_paths_of_the_trace_kernel):
# built here so that the GPU box does not spend its time in nvcc. Never loaded by the package itself.
mkdir -p "$here/test_builds"
for flag in RTX_MBOX_CAP=1 RTX_QUEUE_CAP=1; do
    "$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -D$flag \
        -I"$here/../include" -I"$here/csrc" -shared \
        "$here/csrc/api.cu" "$here/csrc/trace.cu" "$here/csrc/trace_grid.cu" "$here/csrc/aux_kernels.cu" "$here/csrc/tonemap.cu" \
        -o "$here/test_builds/librtx_b200_$flag.so" &
done
wait
cat $(ls "$here"/csrc/*.cu | LC_ALL=C sort) "$here/csrc/rtx_device.cuh" "$here/csrc/trace_common.cuh" "$here/../include/rtx_b200.h" \
    | sha256sum | cut -d' ' -f1 > "$here/test_builds/SOURCES.sha256"
echo "built $here/test_builds/"
# C++ host facade example: the reference's main loop, headless (writes PPM). Links the C ABI only.
g++ -std=c++17 -O2 -ffp-contract=off -Wall -I"$here/../include" -I"$here/host" \
    "$here/examples/headless_main.cpp" -o "$here/rtx_headless" \
    -L"$here" -lrtx_b200 -Wl,-rpath,'$ORIGIN'
echo "built $here/rtx_headless"
g++ -std=c++17 -O2 -ffp-contract=off -Wall -I"$here/../include" -I"$here/host" \
    "$here/examples/camera_walk.cpp" -o "$here/rtx_camera_walk" \
    -L"$here" -lrtx_b200 -Wl,-rpath,'$ORIGIN'
echo "built $here/rtx_camera_walk"
