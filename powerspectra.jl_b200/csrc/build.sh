#!/bin/bash
# Builds libpsb200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${PSB200_OUT:-$here/../libpsb200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC -Xcompiler -Wall -shared ${PSB200_NVCC_EXTRA:-} \
    -ccbin /usr/bin/g++ \
    -o "$out" "$here/psb200.cu" -lcudart_static -lpthread -ldl -lrt
echo "built $out"
