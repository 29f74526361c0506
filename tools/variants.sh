#!/usr/bin/env bash
# Developer tool (GPU box): times config C3 with every librtx_b200_<variant>.so found in the package directory.
cd "$(dirname "$0")/.."
for lib in ray-tracer-from-scratch_b200/librtx_b200_*.so; do
    v=$(basename "$lib" .so | sed 's/librtx_b200_//')
    ms=$(RTX_B200_LIB="$PWD/$lib" timeout 120 python tools/probe.py c3 2>&1 | grep "^c3" | tail -1 | sed -E "s/.*'raytracing_ms': ([0-9.]+).*/\1/")
    echo "$v raytracing_ms=$ms"
done
