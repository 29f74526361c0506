#!/usr/bin/env bash
# Developer tool: builds ray-tracer-from-scratch_b200/librtx_b200_<name>.so with extra -D flags, e.g.
#   tools/build_variant.sh tail RTX_TAIL_TRACE=1
#   tools/build_variant.sh p8 RTX_PAIRS=8
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/../ray-tracer-from-scratch_b200" && pwd)"
name="$1"; shift
defs=""
for d in "$@"; do defs="$defs -D$d"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off $defs \
    -I"$here/../include" -I"$here/csrc" -shared \
    "$here/csrc/api.cu" "$here/csrc/trace.cu" "$here/csrc/trace_grid.cu" "$here/csrc/aux_kernels.cu" "$here/csrc/tonemap.cu" \
    -o "$here/librtx_b200_$name.so"
echo "built $here/librtx_b200_$name.so"
