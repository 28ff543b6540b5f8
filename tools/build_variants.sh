#!/bin/bash
# Builds tiling variants of libpsb200.so into tools/_build/ (development A/B; the product build is csrc/build.sh).
#   tools/build_variants.sh name "<nvcc -D flags>" [name "<flags>" ...]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$here/_build"
while [ $# -ge 2 ]; do
    PSB200_OUT="$here/_build/libpsb200_$1.so" PSB200_NVCC_EXTRA="$2" bash "$here/../powerspectra.jl_b200/csrc/build.sh" > "$here/_build/$1.log" 2>&1 &
    shift 2
    while [ "$(jobs -r | wc -l)" -ge 4 ]; do sleep 1; done
done
wait
ls -la "$here/_build/"*.so
