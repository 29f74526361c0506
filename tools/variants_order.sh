#!/usr/bin/env bash
# Developer tool (GPU box): scan-order and cost-order frame times of every librtx_b200_<variant>.so (tools/order_probe.py).
cd "$(dirname "$0")/.."
for pass in 1 2; do
for lib in ray-tracer-from-scratch_b200/librtx_b200_*.so; do
    v=$(basename "$lib" .so | sed 's/librtx_b200_//')
    RTX_B200_LIB="$PWD/$lib" PROBE_QUICK=1 timeout 120 python tools/order_probe.py ${@:-3840} 2>&1 | sed "s/^/$v  /"
done
done
