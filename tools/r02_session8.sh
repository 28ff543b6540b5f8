#!/bin/bash
# Round 2, call 8: embedded product pass (v4): variants, whole GPU suite, bench line, ncu of the default build.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
step() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
: > gpurun_out/r02_s8_probe.jsonl
for v in embed noembed embl4 embrb8 embl4r8; do
  step "probe $v"
  PSB200_LIB=$PWD/tools/_build/libpsb200_$v.so timeout 120 python tools/kernel_probe.py $v >> gpurun_out/r02_s8_probe.jsonl 2> gpurun_out/r02_s8_probe_$v.err || echo "probe $v failed"
  tail -1 gpurun_out/r02_s8_probe.jsonl | cut -c1-260
done
step "gpu suite (default build = embed)"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_s8_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -6 gpurun_out/r02_s8_gpu_tests.log
step "bench"
timeout 500 python bench.py > gpurun_out/r02_s8_bench.json 2> gpurun_out/r02_s8_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_s8_bench.err; tail -1 gpurun_out/r02_s8_bench.json | cut -c1-600
step "ncu embed"
PROBE_ONCE=1 PSB200_LIB=$PWD/tools/_build/libpsb200_embed.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_kernel_v4 -f -o gpurun_out/r02_ncu_v4_embed python tools/kernel_probe.py embed > gpurun_out/r02_s8_ncu.log 2>&1; echo "ncu rc=$?"
step "done"
