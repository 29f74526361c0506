#!/usr/bin/env bash
# Builds oracle/_ref/libref_oracle.so from the UNMODIFIED reference sources where they lie
# (default /root/reference) plus oracle/ref_harness.cpp and the headless SDL stub.
# TEST INFRASTRUCTURE ONLY. Flags follow the reference build (CMakeLists.txt:5,23: C++17, -O3; no
# -march, no -ffast-math) plus -fopenmp for the row-parallel harness loop (README.md:13) and
# -Dmain=ref_main so main.cpp's main() becomes a callable symbol.
# No reference source is copied: only the compiled .so lands in oracle/_ref/ (git-ignored, but it
# travels to the GPU box with the repo snapshot).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${RTX_REFERENCE_DIR:-/root/reference}"
out="$here/_ref"
if [ ! -f "$ref/main.cpp" ]; then
    echo "build_ref.sh: reference sources not found under $ref (expected on the GPU box; using prebuilt .so if any)" >&2
    exit 3
fi
mkdir -p "$out"
g++ -std=c++17 -O3 -fopenmp -fPIC -shared -w \
    -I"$here/stub" -I"$ref" -I"$here/../include" \
    -Dmain=ref_main \
    "$here/ref_harness.cpp" "$ref/main.cpp" "$ref/vec.cpp" "$ref/scene.cpp" \
    -o "$out/libref_oracle.so"
echo "built $out/libref_oracle.so"
