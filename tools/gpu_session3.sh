#!/bin/bash
# A/B of QuickPol kernel builds (register caps / unroll) on one GPU; parity tests on the candidates.
# The variant libraries are built beforehand in the build container, e.g.
#   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DPSB200_QP_MINBLOCKS=14 \
#        -ccbin /usr/bin/g++ -o tools/_build/libpsb200_mb14.so powerspectra.jl_b200/csrc/psb200.cu -lcudart_static -lpthread -ldl -lrt
# (mb12/14/16: -DPSB200_QP_MINBLOCKS=..; un8: -DPSB200_QP_UNROLL=8) and selected with PSB200_LIB.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/qp_ab.jsonl
for lib in default tools/_build/libpsb200_mb12.so tools/_build/libpsb200_mb14.so tools/_build/libpsb200_mb16.so tools/_build/libpsb200_un8.so tools/_build/libpsb200_un8mb16.so; do
  if [ "$lib" = default ]; then unset PSB200_LIB; else export PSB200_LIB="$PWD/$lib"; fi
  QP_PROBE_FAST=1 QP_PROBE_VARIANTS=tab,simple timeout 40 python tests/tools/quickpol_probe.py 6143 128 2>&1 | tail -1 >> gpurun_out/qp_ab.jsonl
done
cat gpurun_out/qp_ab.jsonl | cut -c1-400
for lib in tools/_build/libpsb200_mb14.so tools/_build/libpsb200_mb16.so; do
  PSB200_LIB="$PWD/$lib" timeout 60 python -m pytest tests/test_quickpol.py -m gpu -x -q 2>&1 | tail -1
done
unset PSB200_LIB
timeout 60 python -m pytest tests/test_quickpol.py -m gpu -x -q 2>&1 | tail -1
