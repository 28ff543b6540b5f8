#!/bin/bash
# Round 2, call 6: v4 with power-of-two rings: tiling variants (rows per warp 2/4/8, pairs per thread), parity of the default build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
: > gpurun_out/r02_s6_probe.jsonl
for v in base2 nr4 nr8 nr4rh8 rl12 nr4rl12 nr4rb8 nr4minb8; do
  step "probe $v"
  PSB200_LIB=$PWD/tools/_build/libpsb200_$v.so timeout 120 python tools/kernel_probe.py $v >> gpurun_out/r02_s6_probe.jsonl 2> gpurun_out/r02_s6_probe_$v.err || echo "probe $v failed"
  tail -1 gpurun_out/r02_s6_probe.jsonl | cut -c1-260
done
step "gpu parity tests (default build)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_s6_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -5 gpurun_out/r02_s6_gpu_tests.log
step "done"
