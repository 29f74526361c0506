#!/usr/bin/env bash
# Developer tool (GPU box): times workloads with every librtx_b200_<variant>.so found in the package directory.
#   tools/variants.sh [c3] [c3small] ...
cd "$(dirname "$0")/.."
for lib in ray-tracer-from-scratch_b200/librtx_b200_*.so; do
    v=$(basename "$lib" .so | sed 's/librtx_b200_//')
    for w in ${@:-c3}; do
        out=$(RTX_B200_LIB="$PWD/$lib" timeout 120 python tools/probe.py $w 2>&1 | grep "^c[0-9]" | tail -1 | sed -E "s/.*'raytracing_ms': ([0-9.]+).*'drain_ms': ([0-9.]+), 'exit_spread_ms': ([0-9.]+).*/raytracing_ms=\1 drain_ms=\2/")
        echo "$v $w $out"
    done
done
